// sp_long_word.cuh — exact BPE merge of one arbitrarily long pre-token by one warp, in global memory.
//
// The lane-per-word and cooperative paths of sp_encode.cu keep a word's symbols in shared memory
// (<= 512 chars).  A longer whitespace-free run (base64, minified code, CJK text under a
// whitespace-split vocabulary) is streamed into a slot of a small global scratch pool instead and
// merged there with the same rule as bpe_model.cc Model::Encode: always the best-priority adjacent
// pair, leftmost on ties.  Structure: doubly linked symbols + per-32-symbol block minima of
// (priority << 32 | position), so one merge costs O(n / 1024) warp steps instead of O(n / 32).
#pragma once
#include <stdint.h>

namespace xllm {

struct LongSlot {  // views into one slot of the pool
  uint32_t* sym;
  uint32_t* prio;
  uint32_t* merged;
  uint32_t* next;
  uint32_t* prev;
  unsigned long long* bmin;
};

__host__ __device__ inline size_t long_slot_bytes(uint32_t cap) {
  return (size_t)cap * 5 * sizeof(uint32_t) + (size_t)(cap / 32 + 1) * sizeof(unsigned long long);
}

__device__ __forceinline__ LongSlot long_slot_view(uint8_t* pool, uint32_t cap, int slot) {
  uint8_t* b = pool + (size_t)slot * long_slot_bytes(cap);
  LongSlot s;
  s.sym = reinterpret_cast<uint32_t*>(b);
  s.prio = s.sym + cap;
  s.merged = s.prio + cap;
  s.next = s.merged + cap;
  s.prev = s.next + cap;
  s.bmin = reinterpret_cast<unsigned long long*>(s.prev + cap);
  return s;
}

// Spins until a slot is free (holders always make progress on their own, so this cannot deadlock).
__device__ __forceinline__ int long_slot_acquire(int* locks, int n_slots, int lane) {
  int got = -1;
  if (lane == 0) {
    unsigned int start = (blockIdx.x * 2654435761u) % (unsigned)n_slots;
    for (;;) {
      for (int k = 0; k < n_slots && got < 0; ++k) {
        const int s = (int)((start + k) % (unsigned)n_slots);
        if (atomicCAS(&locks[s], 0, 1) == 0) got = s;
      }
      if (got >= 0) break;
      __nanosleep(2000);
    }
    __threadfence();
  }
  return __shfl_sync(0xffffffffu, got, 0);
}
__device__ __forceinline__ void long_slot_release(int* locks, int slot, int lane) {
  __syncwarp();
  if (lane == 0) {
    __threadfence();
    atomicExch(&locks[slot], 0);
  }
}

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t < v ? t : v;
  }
  return v;
}

// Recomputes the minimum of block b (positions 32b .. 32b+31, dead / last entries carry kNoPrio).
__device__ __forceinline__ void long_block_min(const LongSlot& L, uint32_t n, uint32_t b, int lane) {
  const uint32_t j = b * 32 + lane;
  unsigned long long key = ~0ull;
  if (j < n) key = ((unsigned long long)L.prio[j] << 32) | j;
  key = warp_min_u64(key);
  if (lane == 0) L.bmin[b] = key;
  __syncwarp();
}

}  // namespace xllm
