// TEST INFRASTRUCTURE.  <xxhash.h> for the reference build = upstream xxHash v0.8.3, header-only, as vendored in
// this image by pyarrow (the build script passes its directory with -I as XLLM_XXHASH_DIR).  XXH3 output has been
// frozen since v0.8.0, so this is the same function the reference's pinned submodule computes.
#pragma once
#ifndef XXH_INLINE_ALL
#define XXH_INLINE_ALL
#endif
#include "arrow/vendored/xxhash/xxhash.h"
