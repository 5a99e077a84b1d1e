// TEST INFRASTRUCTURE.  common/global_gflags.cpp registers three flag validators through this brpc header; flags are
// plain globals in the _ref build, so the registration is a no-op.
#pragma once
#define BRPC_VALIDATE_GFLAG(flag, validate_fn)
