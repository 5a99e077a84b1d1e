// prefix_index.cu — see prefix_index.cuh for the reference mapping.
#include "prefix_index.cuh"

#include <string.h>

#include "common.cuh"

namespace xllm {

namespace {

constexpr uint32_t kEmpty = 0, kFullSlot = 1, kTomb = 2, kBusy = 3;

__device__ __forceinline__ uint64_t home_of(uint64_t lo, uint64_t hi) {
  // keys are XXH3-128 outputs: already uniformly distributed; fold both halves
  return lo ^ (hi >> 17) ^ (hi << 29);
}
inline uint64_t rd64h(const uint8_t* p) {
  uint64_t v;
  memcpy(&v, p, 8);
  return v;
}

// Returns the slot index holding (lo, hi) or ~0 when absent.
__device__ __forceinline__ uint64_t find_slot(const IndexSlot* __restrict__ slots, uint64_t mask, uint64_t lo,
                                              uint64_t hi) {
  uint64_t s = home_of(lo, hi) & mask;
  for (uint64_t probes = 0; probes <= mask; ++probes) {
    const ulonglong2 k = *reinterpret_cast<const ulonglong2*>(&slots[s].klo);
    const uint32_t st = slots[s].state;
    if (st == kEmpty) return ~0ull;
    if (st == kFullSlot && k.x == lo && k.y == hi) return s;
    s = (s + 1) & mask;
  }
  return ~0ull;
}

// ------------------------------------------------------------------ write path
struct StageHeader {  // layout of the staged upload
  int64_t n_keys;
};

// Phase A: one thread per staged key.  Replays the key's ops against its current entry with the exact
// rules of record_updated_kvcaches (global_kvcache_mgr.cpp:177-225) / upload_kvcache (:227-247) /
// update_kvcache (:133-175); updates or tombstones in place; queues keys that must be inserted.
__global__ void index_apply_kernel(IndexSlot* __restrict__ slots, uint64_t mask, const uint64_t* __restrict__ keys,
                                   const int64_t* __restrict__ op_off, const uint64_t* __restrict__ ops,
                                   const uint64_t* __restrict__ payload, int64_t n_keys,
                                   uint64_t* __restrict__ insert_list, int64_t* __restrict__ counters) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_keys) return;
  const uint64_t lo = keys[2 * k], hi = keys[2 * k + 1];
  const uint64_t s = find_slot(slots, mask, lo, hi);
  const bool found = s != ~0ull;
  bool present = found, staged = false;
  uint64_t hbm = 0, dram = 0, ssd = 0;
  if (found) { hbm = slots[s].hbm; dram = slots[s].dram; ssd = slots[s].ssd; }
  for (int64_t i = op_off[k]; i < op_off[k + 1]; ++i) {
    const uint64_t op = ops[i];   // type | instance << 8 | payload index << 32
    const uint32_t type = (uint32_t)op & 0xFF;
    const uint64_t bit = 1ull << ((op >> 8) & 63);
    if (type == 0) {  // stored
      if (!staged) { staged = true; if (!present) { present = true; hbm = dram = ssd = 0; } }
      hbm |= bit;
    } else if (type == 1) {  // offload: HBM -> DRAM, otherwise (DRAM ->) SSD
      if (!staged) { if (!present) continue; staged = true; }
      if (hbm & bit) { hbm &= ~bit; dram |= bit; }
      else { dram &= ~bit; ssd |= bit; }
    } else if (type == 2) {  // removed
      if (!staged) { if (!present) continue; staged = true; }
      hbm &= ~bit; dram &= ~bit; ssd &= ~bit;
    } else if (type == 3) {  // replica PUT: insert_or_assign
      const uint32_t p = (uint32_t)(op >> 32);
      present = true;
      hbm = payload[3 * (size_t)p]; dram = payload[3 * (size_t)p + 1]; ssd = payload[3 * (size_t)p + 2];
    } else {  // replica DELETE
      present = false; staged = false;
      hbm = dram = ssd = 0;
    }
  }
  if (staged && (hbm | dram | ssd) == 0) present = false;  // upload_kvcache erases empty entries
  if (found) {
    if (present) { slots[s].hbm = hbm; slots[s].dram = dram; slots[s].ssd = ssd; }
    else {
      slots[s].state = kTomb;
      atomicAdd((unsigned long long*)&counters[0], (unsigned long long)-1ll);
      atomicAdd((unsigned long long*)&counters[3], 1ull);
    }
  } else if (present) {
    const unsigned long long at = atomicAdd((unsigned long long*)&counters[1], 1ull);
    insert_list[5 * at + 0] = lo; insert_list[5 * at + 1] = hi;
    insert_list[5 * at + 2] = hbm; insert_list[5 * at + 3] = dram; insert_list[5 * at + 4] = ssd;
  }
}

// Phase B: the queued keys are distinct and absent, so each only has to claim the first free slot of
// its probe sequence (no key comparison => no read/write race with other inserters).
__global__ void index_insert_kernel(IndexSlot* __restrict__ slots, uint64_t mask,
                                    const uint64_t* __restrict__ insert_list, int64_t* __restrict__ counters,
                                    int64_t capacity) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= counters[1]) return;
  const uint64_t lo = insert_list[5 * i], hi = insert_list[5 * i + 1];
  const long long live = (long long)atomicAdd((unsigned long long*)&counters[0], 1ull);
  if (live >= capacity) {  // over the configured key capacity: refuse, flag
    atomicAdd((unsigned long long*)&counters[0], (unsigned long long)-1ll);
    atomicExch((unsigned long long*)&counters[2], 1ull);
    return;
  }
  uint64_t s = home_of(lo, hi) & mask;
  for (uint64_t probes = 0; probes <= mask; ++probes) {
    const uint32_t st = slots[s].state;
    if ((st == kEmpty || st == kTomb) && atomicCAS(&slots[s].state, st, kBusy) == st) {
      // claimed: a busy slot is skipped by find_slot, so the key only becomes visible complete
      slots[s].klo = lo; slots[s].khi = hi;
      slots[s].hbm = insert_list[5 * i + 2]; slots[s].dram = insert_list[5 * i + 3]; slots[s].ssd = insert_list[5 * i + 4];
      __threadfence();
      atomicExch(&slots[s].state, kFullSlot);
      if (st == kTomb) atomicAdd((unsigned long long*)&counters[3], (unsigned long long)-1ll);
      return;
    }
    s = (s + 1) & mask;
  }
  atomicExch((unsigned long long*)&counters[2], 1ull);
}

// ------------------------------------------------------------------ read path
__global__ void index_probe_kernel(const IndexSlot* __restrict__ slots, uint64_t mask,
                                   const uint64_t* __restrict__ keys, int64_t n_keys,
                                   uint64_t* __restrict__ masks3) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_keys) return;
  const ulonglong2 key = *reinterpret_cast<const ulonglong2*>(keys + 2 * k);
  const uint64_t s = find_slot(slots, mask, key.x, key.y);
  uint64_t h = 0, d = 0, v = 0;
  if (s != ~0ull) { h = slots[s].hbm; d = slots[s].dram; v = slots[s].ssd; }
  masks3[3 * k] = h; masks3[3 * k + 1] = d; masks3[3 * k + 2] = v;
}

// GlobalKVCacheMgr::match's scan + CacheAwareRouting, one warp per request, lane = block.
//   kProbe = true  (replicated index): the lane probes its block's key itself — keys in, decision out, one kernel;
//                  waves of 32 blocks, and the scan stops after the wave that holds the first miss, so a request is
//                  never probed past it (the reference's loop breaks there, global_kvcache_mgr.cpp:127-129).
//   kProbe = false (hash-range-sharded index): the tier masks were probed on the owning GPUs and came back through
//                  the exchange; the lane reads its block's three masks from masks3.
template <bool kProbe>
__global__ void __launch_bounds__(128) match_route_kernel(const IndexSlot* __restrict__ slots, uint64_t slot_mask,
                                                          const uint64_t* __restrict__ keys,
                                                          const uint64_t* __restrict__ masks3,
                                                          const uint32_t* __restrict__ pos,
                                                          const int64_t* __restrict__ key_start,
                                                          const int32_t* __restrict__ n_blocks, int n_req,
                                                          const InstanceTable* __restrict__ inst,
                                                          MatchOut* __restrict__ match,
                                                          RoutingOut* __restrict__ routing) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n_req) return;
  const int nb = n_blocks[r];
  const int64_t k0 = key_start[r];
  // lane i owns instances i and i + 32: score = 1 + last matched block index holding the instance
  uint32_t sc[3][2] = {{0, 0}, {0, 0}, {0, 0}};
  uint64_t inst_mask = 0;
  int matched = 0;
  bool stop = false;
  for (int base = 0; base < nb && !stop; base += 32) {
    const int i = base + lane;
    uint64_t t[3] = {0, 0, 0};
    if (i < nb) {
      if (kProbe) {
        const ulonglong2 key = *reinterpret_cast<const ulonglong2*>(keys + 2 * (k0 + i));
        const uint64_t s = find_slot(slots, slot_mask, key.x, key.y);
        if (s != ~0ull) {
          const ulonglong2 ds = *reinterpret_cast<const ulonglong2*>(&slots[s].dram);   // the slot's second sector
          t[0] = slots[s].hbm; t[1] = ds.x; t[2] = ds.y;
        }
      } else {
        // sharded: the masks came back in the order the keys were sent; pos maps key index -> that slot
        const size_t at = pos ? (size_t)pos[k0 + i] : (size_t)(k0 + i);
        if (at != (size_t)0xFFFFFFFFu) {
          const uint64_t* m = masks3 + 3 * at;
          t[0] = m[0]; t[1] = m[1]; t[2] = m[2];
        }
      }
    }
    const bool hit = i < nb && (t[0] | t[1] | t[2]) != 0;  // absent or empty entry => miss (:96,127-129)
    const uint32_t hits = __ballot_sync(0xffffffffu, hit);
    const int n_here = nb - base < 32 ? nb - base : 32;
    const uint32_t valid = n_here == 32 ? 0xffffffffu : ((1u << n_here) - 1);
    const uint32_t miss = ~hits & valid;
    const int upto = miss ? __ffs(miss) - 1 : n_here;  // blocks of this chunk before the first miss
    stop = miss != 0;
    matched += upto;
    for (int j = 0; j < upto; ++j) {
#pragma unroll
      for (int tier = 0; tier < 3; ++tier) {
        const uint64_t tm = __shfl_sync(0xffffffffu, t[tier], j);
        inst_mask |= tm;
        if ((tm >> lane) & 1ull) sc[tier][0] = (uint32_t)(base + j + 1);
        if ((tm >> (lane + 32)) & 1ull) sc[tier][1] = (uint32_t)(base + j + 1);
      }
    }
  }
  if (match) {
    MatchOut& o = match[r];
    if (lane == 0) {
      o.max_block_num = (uint32_t)nb;
      o.max_matched_block_num = (uint32_t)matched;
      o.instances = inst_mask;
    }
    o.hbm_score[lane] = (uint16_t)sc[0][0]; o.hbm_score[lane + 32] = (uint16_t)sc[0][1];
    o.dram_score[lane] = (uint16_t)sc[1][0]; o.dram_score[lane + 32] = (uint16_t)sc[1][1];
    o.ssd_score[lane] = (uint16_t)sc[2][0]; o.ssd_score[lane + 32] = (uint16_t)sc[2][1];
  }
  if (!routing) return;

  // ---- InstanceMgr::get_load_metrics (instance_mgr.cpp:287-359)
  const uint64_t usable = inst->has_metrics & inst->schedulable;
  const uint64_t cand = inst_mask & usable;
  uint64_t side[2] = {cand & ~inst->decode_type, cand & inst->decode_type};  // prefill-side, decode-side
  uint64_t max_wait[2] = {0, 0};
#pragma unroll
  for (int sd = 0; sd < 2; ++sd) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int id = lane + 32 * half;
      uint64_t w = ((side[sd] >> id) & 1ull) ? inst->waiting[id] : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const uint64_t x = __shfl_xor_sync(0xffffffffu, w, o);
        w = x > w ? x : w;
      }
      max_wait[sd] = w > max_wait[sd] ? w : max_wait[sd];
    }
  }
  if (side[0] == 0 || side[1] == 0) {
    // fallback: the schedulable instance of that side with the least gpu_cache_usage_perc (< 1, strict);
    // ties resolve to the lowest id here (the reference: unordered_map iteration order)
#pragma unroll
    for (int sd = 0; sd < 2; ++sd) {
      if (side[sd] != 0) continue;
      const uint64_t pool = sd == 0 ? (usable & ~inst->decode_type) : (usable & inst->decode_type);
      float best = 1.0f;
      int best_id = -1;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int id = lane + 32 * half;
        if ((pool >> id) & 1ull) {
          const float u = inst->usage[id];
          if (u < best) { best = u; best_id = id; }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best_id, o);
        if (oi >= 0 && (best_id < 0 || ob < best || (ob == best && oi < best_id))) { best = ob; best_id = oi; }
      }
      if (best_id >= 0) side[sd] = 1ull << best_id;  // max_wait[sd] stays 0 (never updated for the fallback)
    }
  }
  RoutingOut out;
  out.prefill_id = out.decode_id = -1;
  out.prefill_score = out.decode_score = -2.0f;
  out.ok = side[0] != 0;
  if (out.ok) {
    // ---- cost_function (cache_aware_routing.cpp:59-85): HBM scores for both sides (:41,49)
#pragma unroll
    for (int sd = 0; sd < 2; ++sd) {
      if (side[sd] == 0) continue;
      float best = -2.0f;
      int best_id = -1;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int id = lane + 32 * half;
        if ((side[sd] >> id) & 1ull) {
          const uint32_t matched_blocks = sc[0][half];
          const uint32_t q1 = nb == 0 ? 0u : matched_blocks / (uint32_t)nb;
          const uint64_t q2 = max_wait[sd] == 0 ? 0ull : inst->waiting[id] / max_wait[sd];
          float score = (float)q1 - inst->usage[id];
          score = score - (float)q2;
          if (score > best) { best = score; best_id = id; }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best_id, o);
        if (oi >= 0 && (best_id < 0 || ob > best || (ob == best && oi < best_id))) { best = ob; best_id = oi; }
      }
      if (sd == 0) { out.prefill_id = best_id; out.prefill_score = best; }
      else { out.decode_id = best_id; out.decode_score = best; }
    }
  }
  if (lane == 0) routing[r] = out;
}

}  // namespace

// ------------------------------------------------------------------------------ host side
PrefixIndex::~PrefixIndex() {
  if (slots_) cudaFree(slots_);
  if (d_counters_) cudaFree(d_counters_);
  if (d_stage_) cudaFree(d_stage_);
}

int PrefixIndex::init(int64_t capacity_keys) {
  if (capacity_keys <= 0) return XLLM_ERR_INVALID_ARG;
  uint64_t n = 1024;
  while (n < (uint64_t)capacity_keys * 2) n <<= 1;
  XLLM_CUDA_TRY(cudaMalloc(&slots_, n * sizeof(IndexSlot)));
  XLLM_CUDA_TRY(cudaMemset(slots_, 0, n * sizeof(IndexSlot)));
  XLLM_CUDA_TRY(cudaMalloc(&d_counters_, 8 * sizeof(int64_t)));
  XLLM_CUDA_TRY(cudaMemset(d_counters_, 0, 8 * sizeof(int64_t)));
  n_slots_ = n;
  capacity_ = capacity_keys;
  return XLLM_OK;
}

void PrefixIndex::record(int instance, const uint8_t* stored, size_t ns, const uint8_t* offload, size_t no,
                         const uint8_t* removed, size_t nr) {
  auto push = [&](const uint8_t* k16, uint8_t type) {
    Key128 k{rd64h(k16), rd64h(k16 + 8)};
    auto it = staged_.find(k);
    if (it == staged_.end()) {
      it = staged_.emplace(k, std::vector<Op>()).first;
      staged_order_.push_back(k);
    }
    it->second.push_back(Op{type, (uint8_t)instance, 0});
  };
  // the reference walks stored, then offload, then removed (global_kvcache_mgr.cpp:182,193,211)
  for (size_t i = 0; i < ns; ++i) push(stored + 16 * i, 0);
  for (size_t i = 0; i < no; ++i) push(offload + 16 * i, 1);
  for (size_t i = 0; i < nr; ++i) push(removed + 16 * i, 2);
}

void PrefixIndex::put(const uint8_t* key16, uint64_t hbm, uint64_t dram, uint64_t ssd) {
  Key128 k{rd64h(key16), rd64h(key16 + 8)};
  auto it = staged_.find(k);
  if (it == staged_.end()) {
    it = staged_.emplace(k, std::vector<Op>()).first;
    staged_order_.push_back(k);
  }
  it->second.clear();  // an assignment overrides whatever was staged before it
  it->second.push_back(Op{3, 0, (uint32_t)(payload_.size() / 3)});
  payload_.push_back(hbm);
  payload_.push_back(dram);
  payload_.push_back(ssd);
}

void PrefixIndex::erase(const uint8_t* key16) {
  Key128 k{rd64h(key16), rd64h(key16 + 8)};
  auto it = staged_.find(k);
  if (it == staged_.end()) {
    it = staged_.emplace(k, std::vector<Op>()).first;
    staged_order_.push_back(k);
  }
  it->second.clear();
  it->second.push_back(Op{4, 0, 0});
}

cudaEvent_t PrefixIndex::register_reader() {
  cudaEvent_t ev = nullptr;
  if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> l(ev_mu_);
  reader_events_.push_back(ev);
  return ev;
}
void PrefixIndex::unregister_reader(cudaEvent_t ev) {
  if (!ev) return;
  {
    std::lock_guard<std::mutex> l(ev_mu_);
    for (size_t i = 0; i < reader_events_.size(); ++i)
      if (reader_events_[i] == ev) {
        reader_events_.erase(reader_events_.begin() + (long)i);
        break;
      }
  }
  cudaEventDestroy(ev);
}
// The writer's stream waits for the last read every handle enqueued (an event that was never recorded is
// complete).  New reads cannot be enqueued meanwhile: the caller holds rw_ exclusively.
int PrefixIndex::wait_for_readers(cudaStream_t stream) {
  std::lock_guard<std::mutex> l(ev_mu_);
  for (cudaEvent_t ev : reader_events_) XLLM_CUDA_TRY(cudaStreamWaitEvent(stream, ev, 0));
  return XLLM_OK;
}
int PrefixIndex::read_counters(cudaStream_t stream) {
  int64_t c[4] = {0, 0, 0, 0};
  XLLM_CUDA_TRY(cudaMemcpyAsync(c, d_counters_, sizeof(c), cudaMemcpyDeviceToHost, stream));
  XLLM_CUDA_TRY(cudaStreamSynchronize(stream));
  live_ = c[0];
  tombs_ = c[3];
  return c[2] != 0 ? 1 : 0;
}

// Re-inserts every live slot of `from` into the zeroed table `to` (distinct keys: claim the first empty slot).
__global__ void index_rehash_kernel(const IndexSlot* __restrict__ from, uint64_t n_from, IndexSlot* __restrict__ to,
                                    uint64_t to_mask) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_from) return;
  const IndexSlot src = from[i];
  if (src.state != kFullSlot) return;
  uint64_t s = home_of(src.klo, src.khi) & to_mask;
  for (;;) {
    if (to[s].state == kEmpty && atomicCAS(&to[s].state, kEmpty, kBusy) == kEmpty) {
      to[s].klo = src.klo; to[s].khi = src.khi;
      to[s].hbm = src.hbm; to[s].dram = src.dram; to[s].ssd = src.ssd;
      __threadfence();
      atomicExch(&to[s].state, kFullSlot);
      return;
    }
    s = (s + 1) & to_mask;
  }
}

// Tombstones never turn back into empty slots on their own, so under churn (store, evict, store ...) the empty
// slots that end a miss probe would run out and every miss would scan further and further.  When live keys +
// tombstones pass 70 % of the slots the live entries are re-inserted into a fresh table of the same size (live keys
// alone never exceed 50 %: capacity <= slots / 2), which drops every tombstone.  Needs a second table for the
// duration; runs under the exclusive lock.
int PrefixIndex::rebuild(cudaStream_t stream) {
  IndexSlot* fresh = nullptr;
  if (cudaMalloc(&fresh, n_slots_ * sizeof(IndexSlot)) != cudaSuccess) {
    cudaGetLastError();
    return XLLM_OK;   // no room for the second table now: keep the old one, try again at the next publish
  }
  XLLM_CUDA_TRY(cudaMemsetAsync(fresh, 0, n_slots_ * sizeof(IndexSlot), stream));
  const int threads = 256;
  index_rehash_kernel<<<(unsigned)((n_slots_ + threads - 1) / threads), threads, 0, stream>>>(slots_, n_slots_, fresh,
                                                                                              n_slots_ - 1);
  XLLM_CUDA_TRY(cudaGetLastError());
  XLLM_CUDA_TRY(cudaMemsetAsync(d_counters_ + 3, 0, sizeof(int64_t), stream));
  XLLM_CUDA_TRY(cudaStreamSynchronize(stream));
  cudaFree(slots_);
  slots_ = fresh;
  tombs_ = 0;
  ++rebuilds_;
  return XLLM_OK;
}

int PrefixIndex::publish(cudaStream_t stream) {
  if (!ready()) {
    set_last_error("prefix index not configured (index_capacity == 0)");
    return XLLM_ERR_UNSUPPORTED;
  }
  const int64_t nk = (int64_t)staged_order_.size();
  if (nk == 0) return XLLM_OK;
  // whatever happens below, the staged window is consumed: a failed publish must not poison the next one
  struct ClearStaging {
    PrefixIndex* p;
    ~ClearStaging() { p->staged_.clear(); p->staged_order_.clear(); p->payload_.clear(); }
  } clear_staging{this};
  // pack: keys[2*nk] | op_off[nk+1] | ops[n_ops] (u64: type | instance << 8 | payload index << 32) | payload |
  // insert_list[5*nk]
  size_t n_ops = 0;
  for (const auto& k : staged_order_) n_ops += staged_[k].size();
  std::vector<uint64_t> keys(2 * (size_t)nk);
  std::vector<int64_t> op_off((size_t)nk + 1);
  std::vector<uint64_t> ops(n_ops + 1);
  size_t at = 0;
  for (int64_t i = 0; i < nk; ++i) {
    const Key128& k = staged_order_[(size_t)i];
    keys[2 * (size_t)i] = k.lo;
    keys[2 * (size_t)i + 1] = k.hi;
    op_off[(size_t)i] = (int64_t)at;
    for (const Op& op : staged_[k])
      ops[at++] = (uint64_t)op.type | ((uint64_t)op.instance << 8) | ((uint64_t)op.payload << 32);
  }
  op_off[(size_t)nk] = (int64_t)at;
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t b_keys = al(keys.size() * 8), b_off = al(op_off.size() * 8), b_ops = al(ops.size() * 8),
               b_pay = al(payload_.size() * 8 + 8), b_ins = al((size_t)nk * 40 + 8);
  const size_t total = b_keys + b_off + b_ops + b_pay + b_ins;
  if (total > d_stage_cap_) {
    if (d_stage_) cudaFree(d_stage_);
    d_stage_ = nullptr;
    d_stage_cap_ = 0;
    XLLM_CUDA_TRY(cudaMalloc(&d_stage_, total + total / 2));
    d_stage_cap_ = total + total / 2;
  }
  uint8_t* base = static_cast<uint8_t*>(d_stage_);
  uint64_t* d_keys = reinterpret_cast<uint64_t*>(base);
  int64_t* d_off = reinterpret_cast<int64_t*>(base + b_keys);
  uint64_t* d_ops = reinterpret_cast<uint64_t*>(base + b_keys + b_off);
  uint64_t* d_pay = reinterpret_cast<uint64_t*>(base + b_keys + b_off + b_ops);
  uint64_t* d_ins = reinterpret_cast<uint64_t*>(base + b_keys + b_off + b_ops + b_pay);
  // synchronous copies from pageable vectors: publish is the (3 s) control path, not the request path
  XLLM_CUDA_TRY(cudaMemcpyAsync(d_keys, keys.data(), keys.size() * 8, cudaMemcpyHostToDevice, stream));
  XLLM_CUDA_TRY(cudaMemcpyAsync(d_off, op_off.data(), op_off.size() * 8, cudaMemcpyHostToDevice, stream));
  XLLM_CUDA_TRY(cudaMemcpyAsync(d_ops, ops.data(), ops.size() * 8, cudaMemcpyHostToDevice, stream));
  if (!payload_.empty())
    XLLM_CUDA_TRY(cudaMemcpyAsync(d_pay, payload_.data(), payload_.size() * 8, cudaMemcpyHostToDevice, stream));
  // ---- the table changes from here: no reader may be enqueued or in flight
  std::unique_lock<std::shared_mutex> writer(rw_);
  XLLM_TRY_RC(wait_for_readers(stream));
  XLLM_CUDA_TRY(cudaMemsetAsync(d_counters_ + 1, 0, 2 * sizeof(int64_t), stream));
  const int threads = 128;
  const int grid = (int)((nk + threads - 1) / threads);
  index_apply_kernel<<<grid, threads, 0, stream>>>(slots_, n_slots_ - 1, d_keys, d_off, d_ops, d_pay, nk, d_ins,
                                                   d_counters_);
  XLLM_CUDA_TRY(cudaGetLastError());
  index_insert_kernel<<<grid, threads, 0, stream>>>(slots_, n_slots_ - 1, d_ins, d_counters_, capacity_);
  XLLM_CUDA_TRY(cudaGetLastError());
  const int overflow = read_counters(stream);   // synchronises
  if (overflow < 0) return overflow;
  if ((live_ + tombs_) * 10 > (int64_t)n_slots_ * 7) XLLM_TRY_RC(rebuild(stream));
  if (overflow) {
    set_last_error("prefix index is full (capacity %lld keys): some stored keys were dropped", (long long)capacity_);
    return XLLM_ERR_CAPACITY;
  }
  return XLLM_OK;
}

// One thread per slot: drop instance `bit` from the three tiers; an entry left empty is erased (what the
// reference's removed_cache events for every block of a departed instance would add up to).
__global__ void index_clear_instance_kernel(IndexSlot* __restrict__ slots, uint64_t n_slots, uint64_t bit,
                                            int64_t* __restrict__ counters) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_slots || slots[i].state != kFullSlot) return;
  const uint64_t h = slots[i].hbm, d = slots[i].dram, v = slots[i].ssd;
  if (((h | d | v) & bit) == 0) return;
  slots[i].hbm = h & ~bit; slots[i].dram = d & ~bit; slots[i].ssd = v & ~bit;
  if (((h | d | v) & ~bit) == 0) {
    slots[i].state = kTomb;
    atomicAdd((unsigned long long*)&counters[0], (unsigned long long)-1ll);
    atomicAdd((unsigned long long*)&counters[3], 1ull);
  }
}

int PrefixIndex::clear_instance(cudaStream_t stream, int id) {
  if (!ready()) return XLLM_ERR_UNSUPPORTED;
  if (id < 0 || id >= kMaxInstances) return XLLM_ERR_INVALID_ARG;
  std::unique_lock<std::shared_mutex> writer(rw_);
  XLLM_TRY_RC(wait_for_readers(stream));
  const int threads = 256;
  index_clear_instance_kernel<<<(unsigned)((n_slots_ + threads - 1) / threads), threads, 0, stream>>>(
      slots_, n_slots_, 1ull << id, d_counters_);
  XLLM_CUDA_TRY(cudaGetLastError());
  const int rc = read_counters(stream);
  if (rc < 0) return rc;
  if ((live_ + tombs_) * 10 > (int64_t)n_slots_ * 7) XLLM_TRY_RC(rebuild(stream));
  return XLLM_OK;
}

int PrefixIndex::size(cudaStream_t stream, int64_t* n) {
  if (!ready()) { *n = 0; return XLLM_OK; }
  std::shared_lock<std::shared_mutex> reader(rw_);
  XLLM_CUDA_TRY(cudaMemcpyAsync(n, d_counters_, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
  XLLM_CUDA_TRY(cudaStreamSynchronize(stream));
  return XLLM_OK;
}

// One thread per slot: live slots are compacted (order = atomic arrival) into a packed array of
// {key lo, key hi, hbm, dram, ssd} rows.
__global__ void index_export_kernel(const IndexSlot* __restrict__ slots, uint64_t n_slots, uint64_t* __restrict__ rows,
                                    long long cap, unsigned long long* __restrict__ counter) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_slots) return;
  const IndexSlot s = slots[i];
  if (s.state != 1) return;
  const unsigned long long at = atomicAdd(counter, 1ull);
  if ((long long)at < cap) {
    uint64_t* r = rows + at * 5;
    r[0] = s.klo; r[1] = s.khi; r[2] = s.hbm; r[3] = s.dram; r[4] = s.ssd;
  }
}

int PrefixIndex::export_all(cudaStream_t stream, int64_t cap, uint8_t* keys16, uint64_t* hbm, uint64_t* dram,
                            uint64_t* ssd, int64_t* n_out) {
  if (!ready()) return XLLM_ERR_UNSUPPORTED;
  std::shared_lock<std::shared_mutex> reader(rw_);   // synchronises before returning: no event needed
  if (cap < 0) cap = 0;
  const size_t need = 64 + (size_t)cap * 40;
  if (d_stage_cap_ < need) {
    if (d_stage_) cudaFree(d_stage_);
    d_stage_ = nullptr;
    d_stage_cap_ = 0;
    if (cudaMalloc(&d_stage_, need) != cudaSuccess) {
      set_last_error("cudaMalloc(%zu) for the index snapshot failed", need);
      return XLLM_ERR_NOMEM;
    }
    d_stage_cap_ = need;
  }
  unsigned long long* d_count = static_cast<unsigned long long*>(d_stage_);
  uint64_t* d_rows = reinterpret_cast<uint64_t*>(static_cast<uint8_t*>(d_stage_) + 64);
  XLLM_CUDA_TRY(cudaMemsetAsync(d_count, 0, 8, stream));
  const int threads = 256;
  index_export_kernel<<<(unsigned)((n_slots_ + threads - 1) / threads), threads, 0, stream>>>(
      slots_, n_slots_, d_rows, (long long)cap, d_count);
  XLLM_CUDA_TRY(cudaGetLastError());
  unsigned long long n = 0;
  XLLM_CUDA_TRY(cudaMemcpyAsync(&n, d_count, 8, cudaMemcpyDeviceToHost, stream));
  XLLM_CUDA_TRY(cudaStreamSynchronize(stream));
  *n_out = (int64_t)n;
  const int64_t take = (int64_t)n < cap ? (int64_t)n : cap;
  if (take > 0) {
    std::vector<uint64_t> rows((size_t)take * 5);
    XLLM_CUDA_TRY(cudaMemcpyAsync(rows.data(), d_rows, rows.size() * 8, cudaMemcpyDeviceToHost, stream));
    XLLM_CUDA_TRY(cudaStreamSynchronize(stream));
    for (int64_t i = 0; i < take; ++i) {
      memcpy(keys16 + 16 * i, &rows[(size_t)i * 5], 16);  // low64 LE || high64 LE, as hashed (hash_util.cpp:26-27)
      hbm[i] = rows[(size_t)i * 5 + 2];
      dram[i] = rows[(size_t)i * 5 + 3];
      ssd[i] = rows[(size_t)i * 5 + 4];
    }
  }
  if ((int64_t)n > cap) {
    set_last_error("index snapshot: %lld live keys, room for %lld", (long long)n, (long long)cap);
    return XLLM_ERR_CAPACITY;
  }
  return XLLM_OK;
}

int PrefixIndex::get(cudaStream_t stream, const uint8_t* key16, uint64_t masks3[3], int* found) {
  if (!ready()) return XLLM_ERR_UNSUPPORTED;
  std::shared_lock<std::shared_mutex> reader(rw_);   // synchronises before returning: no event needed
  if (d_stage_cap_ < 256) {
    if (d_stage_) cudaFree(d_stage_);
    d_stage_ = nullptr;
    XLLM_CUDA_TRY(cudaMalloc(&d_stage_, 4096));
    d_stage_cap_ = 4096;
  }
  uint8_t* base = static_cast<uint8_t*>(d_stage_);
  XLLM_CUDA_TRY(cudaMemcpyAsync(base, key16, 16, cudaMemcpyHostToDevice, stream));
  XLLM_CUDA_TRY(probe(base, 1, reinterpret_cast<uint64_t*>(base + 64), stream));
  XLLM_CUDA_TRY(cudaMemcpyAsync(masks3, base + 64, 24, cudaMemcpyDeviceToHost, stream));
  XLLM_CUDA_TRY(cudaStreamSynchronize(stream));
  *found = (masks3[0] | masks3[1] | masks3[2]) != 0;
  return XLLM_OK;
}

cudaError_t PrefixIndex::probe(const uint8_t* d_keys, int64_t n_keys, uint64_t* d_masks3, cudaStream_t stream) const {
  if (n_keys <= 0) return cudaSuccess;
  const int threads = 256;
  const int64_t grid = (n_keys + threads - 1) / threads;
  index_probe_kernel<<<(unsigned)grid, threads, 0, stream>>>(slots_, n_slots_ - 1,
                                                             reinterpret_cast<const uint64_t*>(d_keys), n_keys,
                                                             d_masks3);
  return cudaGetLastError();
}

cudaError_t PrefixIndex::match_route(const uint8_t* d_keys, const int64_t* d_key_start, const int32_t* d_n_blocks,
                                     int n_req, const InstanceTable* d_instances, MatchOut* d_match,
                                     RoutingOut* d_routing, cudaStream_t stream) const {
  if (n_req <= 0) return cudaSuccess;
  const int warps = 4;
  const int grid = (n_req + warps - 1) / warps;
  match_route_kernel<true><<<grid, warps * 32, 0, stream>>>(slots_, n_slots_ - 1,
                                                            reinterpret_cast<const uint64_t*>(d_keys), nullptr,
                                                            nullptr, d_key_start, d_n_blocks, n_req, d_instances,
                                                            d_match, d_routing);
  return cudaGetLastError();
}

cudaError_t score_route_launch(const uint64_t* d_masks3, const int64_t* d_key_start, const int32_t* d_n_blocks,
                               int n_req, const InstanceTable* d_instances, MatchOut* d_match,
                               RoutingOut* d_routing, cudaStream_t stream, const uint32_t* d_pos) {
  if (n_req <= 0) return cudaSuccess;
  const int warps = 4;
  const int grid = (n_req + warps - 1) / warps;
  match_route_kernel<false><<<grid, warps * 32, 0, stream>>>(nullptr, 0, nullptr, d_masks3, d_pos, d_key_start,
                                                             d_n_blocks, n_req, d_instances, d_match, d_routing);
  return cudaGetLastError();
}

// Owner side of the sharded index: probes the tuples of `world` received messages ([u32 count | 12 bytes | tuples of
// 24 bytes, key first] each, shard_exchange.cuh) and writes {hbm, dram, ssd} per tuple in arrival order.
__global__ void index_probe_messages_kernel(const IndexSlot* __restrict__ slots, uint64_t mask,
                                            const uint8_t* __restrict__ recv, size_t msg_bytes, int world,
                                            uint32_t cap, uint64_t* __restrict__ back) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t p = (uint32_t)(t / cap), j = (uint32_t)(t % cap);
  if ((int)p >= world) return;
  const uint8_t* msg = recv + (size_t)p * msg_bytes;
  uint32_t n = *reinterpret_cast<const uint32_t*>(msg);
  if (n > cap) n = cap;
  if (j >= n) return;
  const uint64_t* key = reinterpret_cast<const uint64_t*>(msg + 16 + (size_t)j * 24);
  const uint64_t s = find_slot(slots, mask, key[0], key[1]);
  uint64_t h = 0, d = 0, v = 0;
  if (s != ~0ull) { h = slots[s].hbm; d = slots[s].dram; v = slots[s].ssd; }
  uint64_t* o = back + ((size_t)p * cap + j) * 3;
  o[0] = h; o[1] = d; o[2] = v;
}

cudaError_t PrefixIndex::probe_messages(const uint8_t* d_recv, size_t msg_bytes, int world, uint32_t cap,
                                        uint64_t* d_back, cudaStream_t stream) const {
  const uint64_t total = (uint64_t)world * cap;
  if (total == 0) return cudaSuccess;
  const int threads = 256;
  index_probe_messages_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, stream>>>(
      slots_, n_slots_ - 1, d_recv, msg_bytes, world, cap, d_back);
  return cudaGetLastError();
}

}  // namespace xllm
