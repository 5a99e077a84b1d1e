// Chained per-block XXH3-128 of token-ID streams — the block-hash step of the
// prefix-cache path.
//
// Replaces, on device, the loop body of GlobalKVCacheMgr::match
// (xllm_service/scheduler/managers/global_kvcache_mgr.cpp:84-94) and
// xxh3_128bits_hash (xllm_service/common/hash_util.cpp:18-45).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <cuda_runtime.h>

namespace xllm {

// Seed-derived constants, computed once on the host (XXH3's "custom secret":
// S[16i..] = LE64(kSecret[16i..]) + seed ; S[16i+8..] = LE64(kSecret[16i+8..]) - seed)
// and passed by value as a __grid_constant__ kernel parameter so every use is a
// constant-bank operand.
struct Xxh3Consts {
  uint64_t s[16];     // custom secret as aligned u64: s[i] = LE64(S + 8i), i = 0..15 (stripe k, lane l uses s[k+l])
  uint64_t last[8];   // LE64(S + 121 + 8l): the last-stripe keys (192 - 64 - 7 = 121)
  uint64_t mlo[8];    // LE64(S + 11 + 8i): mergeAccs keys for low64
  uint64_t mhi[8];    // LE64(S + 117 + 8i): mergeAccs keys for high64 (192 - 64 - 11 = 117)
  uint64_t seed;      // zero-extended FLAGS_xxh3_128bits_seed
  uint8_t secret[192];   // the full custom secret (== kSecret when seed == 0), for the generic path
  uint8_t ksecret[192];  // XXH3_kSecret, used by the <= 240-byte classes
};

void xxh3_make_consts(uint32_t seed, Xxh3Consts* out);

// Row-addressed batch: request r owns tokens[tok_start[r] .. tok_start[r] + n_tok[r]) and
// writes floor(n_tok[r]/block_size) keys of 16 bytes at keys + 16 * key_start[r].
// All pointers are device pointers.  task_counter: one zeroed uint32 in device memory
// (reset by the launcher on `stream`).
cudaError_t xxh3_chain_launch(const int32_t* tokens, const int64_t* tok_start, const int32_t* n_tok,
                              uint8_t* keys, const int64_t* key_start, int n_req, int block_size,
                              const Xxh3Consts& consts, unsigned int* task_counter, cudaStream_t stream);

// One XXH3_128bits_withSeed over `len` device bytes (single thread; the drop-in for a
// lone xxh3_128bits_hash call).  out16 = low64 LE || high64 LE.
cudaError_t xxh3_single_launch(const uint8_t* d_data, size_t len, uint8_t* d_out16, const Xxh3Consts& consts,
                               cudaStream_t stream);

}  // namespace xllm
