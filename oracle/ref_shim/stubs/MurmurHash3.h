// TEST INFRASTRUCTURE.  common/hash_util.cpp includes <MurmurHash3.h> but the hot path never calls it.
#pragma once
