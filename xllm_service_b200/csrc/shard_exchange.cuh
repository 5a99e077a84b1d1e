// shard_exchange.cuh — the hash-range-sharded prefix index (BASELINE config 4).
//
// When the index outgrows one GPU it is split by hash range over the G GPUs of the box: the GPU that owns a block
// key is `low64 >> (64 - log2 G)`.  GlobalKVCacheMgr::match (global_kvcache_mgr.cpp:73-131) for a batch then has
// one real exchange step each way, done natively over NCCL (NVLink / NVSwitch), with no host round trip in between:
//
//   origin GPU   bucket_by_owner_kernel : every (request, block) key -> a 24-byte tuple {hash128, req, blk, src}
//                                         appended to its owner's outgoing message (fixed-capacity slots)
//   NCCL         one grouped ncclSend/ncclRecv round (all-to-all) of [16-byte header | tuples]
//   owner GPU    probe of the received tuples in its slice of the table -> {hbm, dram, ssd} masks, in arrival order
//   NCCL         one grouped round back (24-byte mask triples)
//   origin GPU   match_route_kernel<false>: first-miss scan (absent or empty = miss, :96,127-129), set_score,
//                get_load_metrics, cost_function — through a position map written by the bucketing step
//
// Message capacity is fixed per handle (the same on every rank: it derives from the config), so no count has to
// reach the host before the NCCL calls are issued.  Every header carries the sender's largest bucket and its total
// key count; since every rank hears from every rank, all of them see the same numbers and — in the rare case a
// bucket overflowed (heavily repeated keys all owned by one GPU) — all of them repeat the round with the same larger
// capacity.  NCCL is loaded with dlopen("libnccl.so.2") on first use: handles without a sharded index never need it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>

#include "prefix_index.cuh"

namespace xllm {

struct ShardTuple {   // what travels to the owner: the (hash, request-id) pair of north_star, 24 bytes
  uint64_t lo, hi;    // the 16-byte block key (low64 LE || high64 LE)
  uint32_t req;       // request row on the origin GPU
  uint16_t blk;       // block index inside the request
  uint16_t src;       // origin rank
};
static_assert(sizeof(ShardTuple) == 24, "tuple layout is the wire format");

struct ShardHeader {     // first 16 bytes of every outgoing message
  uint32_t count;        // tuples that follow (<= capacity)
  uint32_t max_bucket;   // the sender's largest bucket this round (may exceed capacity: overflow)
  uint32_t total_keys;   // keys the sender bucketed this round
  uint32_t pad;
};
static_assert(sizeof(ShardHeader) == 16, "header layout is the wire format");

struct ShardTimes {   // device time of the last round, milliseconds (CUDA events on the round's stream)
  float bucket_ms, exchange_out_ms, probe_ms, exchange_back_ms, score_ms;
};

// owner of a key among `world` (a power of two) GPUs: the top log2(world) bits of low64
inline int shard_owner_of(uint64_t lo, int log2_world) { return log2_world == 0 ? 0 : (int)(lo >> (64 - log2_world)); }

class ShardExchange {
 public:
  ~ShardExchange();
  // unique_id: the 128 bytes of an ncclUniqueId (xllm_shard_unique_id), identical on every rank.  Collective.
  int init(int world, int rank, const void* unique_id, int device, int64_t bucket_capacity);
  static int unique_id(void* out128);

  int world() const { return world_; }
  int rank() const { return rank_; }
  int log2_world() const { return log2_; }
  bool owns(uint64_t lo) const { return shard_owner_of(lo, log2_) == rank_; }

  // GlobalKVCacheMgr::match + CacheAwareRouting for n_req requests against the sharded index.  COLLECTIVE: every
  // rank calls it once per batch (with its own requests; n_req may be 0).  Asynchronous on `stream` up to the point
  // where the headers are read back (one small D2H + synchronise at the end, needed to detect an overflow round).
  int match_route(PrefixIndex& index, cudaEvent_t index_read_ev, const uint8_t* d_keys, const int64_t* d_key_start,
                  const int32_t* d_n_blocks, int n_req, int64_t n_keys_bound, const InstanceTable* d_instances,
                  MatchOut* d_match, RoutingOut* d_routing, cudaStream_t stream);
  const ShardTimes& last_times() const { return times_; }
  int64_t bucket_capacity() const { return cap_; }
  int64_t overflow_rounds() const { return overflow_rounds_; }

 private:
  int ensure_buffers(int64_t cap, int64_t n_keys_bound);
  int round(PrefixIndex& index, cudaEvent_t index_read_ev, const uint8_t* d_keys, const int64_t* d_key_start,
            const int32_t* d_n_blocks, int n_req, const InstanceTable* d_instances, MatchOut* d_match,
            RoutingOut* d_routing, cudaStream_t stream, uint32_t* need_cap);
  int exchange(const uint8_t* send, uint8_t* recv, size_t bytes_per_peer, cudaStream_t stream);
  std::mutex mu_;          // one round at a time per communicator (clones share it)
  void* comm_ = nullptr;   // ncclComm_t
  int world_ = 1, rank_ = 0, log2_ = 0, device_ = 0;
  int64_t cap_ = 0;        // tuples per peer message
  int64_t buf_cap_ = 0, pos_cap_ = 0;
  uint8_t *d_send_ = nullptr, *d_recv_ = nullptr;            // [world][16 + cap * 24]
  uint8_t *d_back_send_ = nullptr, *d_back_recv_ = nullptr;  // [world][cap * 24]
  uint32_t* d_cursors_ = nullptr;                            // [world] + overflow flag
  uint32_t* d_pos_ = nullptr;                                // key index -> peer * cap + slot (~0 = overflowed)
  ShardHeader* d_headers_ = nullptr;                         // the world headers received this round, gathered
  ShardHeader* h_headers_ = nullptr;                         // ... and their pinned host copy
  cudaEvent_t ev_[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  ShardTimes times_ = {0, 0, 0, 0, 0};
  int64_t overflow_rounds_ = 0;
};

}  // namespace xllm
