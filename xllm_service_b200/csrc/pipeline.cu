// pipeline.cu — the whole ingest hot path for a batch of requests behind one C-ABI call:
//   host text -> [H2D] -> tokenize -> block-hash chain -> index probe -> match scan + routing -> [D2H]
// i.e. what Scheduler::schedule does per request between scheduler.cpp:128 and :135
// (Tokenizer::encode, then CacheAwareRouting::select_instances_pair -> GlobalKVCacheMgr::match ->
// cost_function), batched.  The batch is cut into chunks that flow through an upload, a kernel and a download
// stream so the PCIe copies of one chunk overlap the kernels of another; the caller's buffers should be
// page-locked (xllm_host_alloc: also NUMA-local to the GPU) for the copies to be asynchronous.
#include <ctype.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "../../include/xllm_ingest.h"
#include <chrono>
#include <vector>

#include "handle.h"
#include "pipeline_schedule.h"

namespace xllm {

namespace {

// Row descriptors for the hash / match stages of one chunk (request-local indices).
__global__ void prep_rows_kernel(const int32_t* __restrict__ n_ids, int n, int64_t ids_stride, int64_t keys_stride,
                                 int block_size, int64_t* __restrict__ tok_start, int32_t* __restrict__ n_tok,
                                 int64_t* __restrict__ key_start, int32_t* __restrict__ n_blocks) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  int64_t t = n_ids[r];
  if (t > ids_stride) t = ids_stride;  // truncated rows hash what was written
  int64_t nb = t / block_size;
  if (nb > keys_stride) nb = keys_stride;
  tok_start[r] = (int64_t)r * ids_stride;
  n_tok[r] = (int32_t)(nb * block_size);
  key_start[r] = (int64_t)r * keys_stride;
  n_blocks[r] = (int32_t)nb;
}

// Sharded index: the exchange runs once per batch, so every chunk also files its rows in batch-wide descriptors.
__global__ void batch_rows_kernel(const int32_t* __restrict__ chunk_n_blocks, int m, int64_t row0, int64_t keys_stride,
                                  int64_t* __restrict__ all_key_start, int32_t* __restrict__ all_n_blocks) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m) return;
  all_key_start[row0 + r] = (row0 + r) * keys_stride;
  all_n_blocks[row0 + r] = chunk_n_blocks[r];
}

}  // namespace

int PipeSlot::ensure(size_t text_bytes, int n, int64_t ids_stride, int64_t keys_stride) {
  int rc;
  if ((rc = d_text.reserve(text_bytes + 64)) != XLLM_OK) return rc;
  if ((rc = d_offsets.reserve((size_t)(n + 1) * 8)) != XLLM_OK) return rc;
  if ((rc = d_ids.reserve((size_t)n * (size_t)ids_stride * 4 + 64)) != XLLM_OK) return rc;
  if ((rc = d_n_ids.reserve((size_t)n * 4)) != XLLM_OK) return rc;
  if ((rc = d_status.reserve((size_t)n * 4)) != XLLM_OK) return rc;
  if ((rc = d_defer.reserve((size_t)n * 4)) != XLLM_OK) return rc;
  if ((rc = d_tok_start.reserve((size_t)n * 8)) != XLLM_OK) return rc;
  if ((rc = d_n_tok.reserve((size_t)n * 4)) != XLLM_OK) return rc;
  if ((rc = d_key_start.reserve((size_t)n * 8)) != XLLM_OK) return rc;
  if ((rc = d_n_blocks.reserve((size_t)n * 4)) != XLLM_OK) return rc;
  if ((rc = d_keys.reserve((size_t)n * (size_t)keys_stride * 16 + 64)) != XLLM_OK) return rc;
  if ((rc = d_match.reserve((size_t)n * sizeof(MatchOut))) != XLLM_OK) return rc;
  if ((rc = d_routing.reserve((size_t)n * sizeof(RoutingOut))) != XLLM_OK) return rc;
  return XLLM_OK;
}

void PipeSlot::release() {
  d_defer.release();
  d_memo.release();
  d_text.release(); d_offsets.release(); d_ids.release(); d_n_ids.release(); d_status.release();
  d_tok_start.release(); d_n_tok.release(); d_key_start.release(); d_n_blocks.release();
  d_keys.release(); d_masks.release(); d_match.release(); d_routing.release();
  for (int k = 0; k < 3; ++k) {
    if (ev[k]) cudaEventDestroy(ev[k]);
    ev[k] = nullptr;
  }
  if (counters) cudaFree(counters);
  counters = nullptr;
  busy = false;
}

}  // namespace xllm

using namespace xllm;

#define XLLM_TRY(expr)              \
  do {                              \
    int _rc = (expr);               \
    if (_rc != XLLM_OK) return _rc; \
  } while (0)

extern "C" {

// CPUs local to the current device's PCIe root (sysfs local_cpulist, e.g. "0-31,64-95")
static bool gpu_local_cpus(cpu_set_t* set) {
  int dev = 0;
  char bus[32] = {0};
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetPCIBusId(bus, (int)sizeof(bus), dev) != cudaSuccess) return false;
  for (char* c = bus; *c; ++c) *c = (char)tolower(*c);
  char path[128];
  snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/local_cpulist", bus);
  FILE* f = fopen(path, "r");
  if (!f) return false;
  char line[512] = {0};
  const bool got = fgets(line, sizeof(line), f) != nullptr;
  fclose(f);
  if (!got) return false;
  CPU_ZERO(set);
  int n = 0;
  for (char* p = line; *p && *p != '\n';) {
    char* end;
    const long a = strtol(p, &end, 10);
    if (end == p) break;
    long b = a;
    p = end;
    if (*p == '-') { b = strtol(p + 1, &end, 10); p = end; }
    for (long c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET((int)c, set); ++n; }
    if (*p == ',') ++p;
  }
  return n > 0;
}

int xllm_host_alloc(void** out, size_t bytes) {
  if (!out) return XLLM_ERR_INVALID_ARG;
  *out = nullptr;
  // Page-locking populates the pages, on the NUMA node of the calling thread: run the allocation on a CPU next to
  // the GPU so the DMA does not cross the socket interconnect (it costs PCIe bandwidth when both directions are busy)
  cpu_set_t old_set, local_set, both;
  bool bound = false;
  if (sched_getaffinity(0, sizeof(old_set), &old_set) == 0 && gpu_local_cpus(&local_set)) {
    CPU_AND(&both, &old_set, &local_set);
    if (CPU_COUNT(&both) > 0 && sched_setaffinity(0, sizeof(both), &both) == 0) bound = true;
  }
  cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
  if (bound) sched_setaffinity(0, sizeof(old_set), &old_set);
  if (e != cudaSuccess) {
    set_last_error("cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return XLLM_ERR_NOMEM;
  }
  return XLLM_OK;
}
void xllm_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int xllm_ingest_batch(xllm_ingest_t h, const xllm_ingest_io* io) {
  if (!h || !io || io->n_req < 0) {
    set_last_error("xllm_ingest_batch: invalid argument");
    return XLLM_ERR_INVALID_ARG;
  }
  const int32_t n = io->n_req;
  if (n == 0) return XLLM_OK;
  if (!io->offsets || !io->n_ids || !io->status || io->ids_stride <= 0 || !io->ids || io->keys_stride < 0 ||
      (io->keys_stride > 0 && !io->keys)) {
    set_last_error("xllm_ingest_batch: missing buffer");
    return XLLM_ERR_INVALID_ARG;
  }
  if (!h->sp_dev) {
    set_last_error("handle has no tokenizer (tokenizer_path was not set)");
    return XLLM_ERR_UNSUPPORTED;
  }
  const bool want_match = io->match != nullptr || io->routing != nullptr;
  if (want_match && (!h->index || !h->index->ready())) {
    set_last_error("match/routing requested but the prefix index is not configured (index_capacity == 0)");
    return XLLM_ERR_UNSUPPORTED;
  }
  const int64_t keys_stride = io->keys_stride > 0 ? io->keys_stride : (want_match ? io->ids_stride / h->block_size : 0);
  for (int32_t r = 0; r < n; ++r)
    if (io->offsets[r + 1] < io->offsets[r] || io->offsets[r] < 0 ||
        io->offsets[r + 1] - io->offsets[r] > 0x7fffffffLL) {
      set_last_error("xllm_ingest_batch: bad offsets at request %d", r);
      return XLLM_ERR_INVALID_ARG;
    }
  if (io->offsets[n] > io->offsets[0] && !io->text) return XLLM_ERR_INVALID_ARG;

  std::lock_guard<std::mutex> lock(h->mu);
  XLLM_CUDA_TRY(cudaSetDevice(h->device));
  // hash-range-sharded index: the chunks tokenise and hash as usual, their keys stay resident, and ONE exchange round
  // for the whole batch follows the last chunk (collective: every rank makes this call once per batch)
  const bool sharded = want_match && h->shard != nullptr;
  if (sharded) {
    XLLM_TRY(h->d_all_keys.reserve((size_t)n * (size_t)keys_stride * 16 + 64));
    XLLM_TRY(h->d_all_key_start.reserve((size_t)n * 8));
    XLLM_TRY(h->d_all_n_blocks.reserve((size_t)n * 4));
    XLLM_TRY(h->d_all_match.reserve((size_t)n * sizeof(MatchOut)));
    XLLM_TRY(h->d_all_routing.reserve((size_t)n * sizeof(RoutingOut)));
  }
  if (want_match) {
    std::lock_guard<std::mutex> l2(*h->index_mu);
    XLLM_CUDA_TRY(cudaMemcpyAsync(h->d_inst, h->inst_host.get(), sizeof(InstanceTable), cudaMemcpyHostToDevice,
                                  h->stream));
    XLLM_CUDA_TRY(cudaStreamSynchronize(h->stream));
  }
  // Three engine streams — uploads, kernels, downloads — joined by events: each hardware engine sees its work in
  // chunk order (no stream-to-queue aliasing, no kernel waiting behind its own chunk's download), and a chunk's
  // buffers (slot = chunk mod pipe_slots) are reused only after the chunk that last held them has left the
  // download engine.
  const int chunk_req = h->pipe_chunk_req;
  const int64_t chunk_bytes = h->pipe_chunk_bytes;
  const int n_slots = h->pipe_slots;
  for (int k = 0; k < 3; ++k)
    if (!h->pipe_stream[k]) XLLM_CUDA_TRY(cudaStreamCreateWithFlags(&h->pipe_stream[k], cudaStreamNonBlocking));
  cudaStream_t s_in = h->pipe_stream[0], s_k = h->pipe_stream[1], s_out = h->pipe_stream[2];
  for (int s = 0; s < n_slots; ++s) {
    PipeSlot& sl = h->pipe[s];
    if (!sl.counters) XLLM_CUDA_TRY(cudaMalloc(&sl.counters, 64));
    for (int k = 0; k < 3; ++k)
      if (!sl.ev[k]) XLLM_CUDA_TRY(cudaEventCreateWithFlags(&sl.ev[k], cudaEventDisableTiming));
    sl.busy = false;
  }
  // XLLM_PIPE_TRACE=1: per-chunk device timeline on stderr (debugging aid)
  static const bool trace = getenv("XLLM_PIPE_TRACE") != nullptr;
  struct TraceRow { int m; double host_ms; cudaEvent_t e[5]; };
  std::vector<TraceRow> rows;
  cudaEvent_t ev_begin = nullptr;
  const auto host_t0 = std::chrono::steady_clock::now();
  if (trace) { cudaEventCreate(&ev_begin); cudaEventRecord(ev_begin, s_in); }
  auto mark = [&](int k, cudaStream_t st) {
    if (!trace) return;
    cudaEventCreate(&rows.back().e[k]);
    cudaEventRecord(rows.back().e[k], st);
  };
  int slot = 0;
  int32_t c0 = 0;
  int rc = XLLM_OK;
  h->last_chunks = 0;
  h->last_launches = 0;
  ChunkSchedule sched(chunk_req);  // ramp-up, full-size bulk, quarter-size tail (pipeline_schedule.h)
  // inside the chunk loop a CUDA error must not return at once: copies into the caller's buffers may be in flight, so
  // leave the loop and synchronise the three streams first
#define PIPE_CUDA_TRY(expr)                                                                              \
  if (cudaError_t _pe = (expr); _pe != cudaSuccess) {                                                    \
    ::xllm::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_pe));       \
    rc = XLLM_ERR_CUDA;                                                                                  \
    break;                                                                                               \
  } else                                                                                                 \
    (void)0
  while (c0 < n) {
    const int64_t target = sched.next((int64_t)n - c0);
    int32_t c1 = c0;
    while (c1 < n && c1 - c0 < target && (c1 == c0 || io->offsets[c1 + 1] - io->offsets[c0] <= chunk_bytes)) ++c1;
    const int m = c1 - c0;
    const int64_t t0 = io->offsets[c0];
    const size_t text_bytes = (size_t)(io->offsets[c1] - t0);
    PipeSlot& sl = h->pipe[slot];
    uint8_t* chunk_keys = sl.d_keys.as<uint8_t>();
    // the slot's previous chunk has been downloaded (host wait: ensure() below may reallocate its buffers)
    if (sl.busy) PIPE_CUDA_TRY(cudaEventSynchronize(sl.ev[2]));
    if ((rc = sl.ensure(text_bytes, m, io->ids_stride, keys_stride)) != XLLM_OK) break;
    if (h->memo_slots && (rc = sl.d_memo.reserve((size_t)h->memo_slots * 32)) != XLLM_OK) break;
    xllm::SpMemo memo;
    memo.table = h->memo_slots ? sl.d_memo.p : nullptr;
    memo.slots = h->memo_slots;
    if (trace) {
      rows.push_back(TraceRow{m, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count(), {}});
      mark(0, s_in);
    }
    // ---- upload
    if (text_bytes)
      PIPE_CUDA_TRY(cudaMemcpyAsync(sl.d_text.p, io->text + t0, text_bytes, cudaMemcpyHostToDevice, s_in));
    PIPE_CUDA_TRY(cudaMemcpyAsync(sl.d_offsets.p, io->offsets + c0, (size_t)(m + 1) * 8, cudaMemcpyHostToDevice, s_in));
    PIPE_CUDA_TRY(cudaEventRecord(sl.ev[0], s_in));
    mark(1, s_in);
    // ---- kernels: tokenize -> row prep -> chained block hash -> index probe -> match scan + routing
    PIPE_CUDA_TRY(cudaStreamWaitEvent(s_k, sl.ev[0], 0));
    PIPE_CUDA_TRY(sp_encode_launch(h->sp_dev->dev(), sl.d_text.as<uint8_t>() - t0, sl.d_offsets.as<int64_t>(), m,
                                   sl.d_ids.as<int32_t>(), io->ids_stride, sl.d_n_ids.as<int32_t>(),
                                   sl.d_status.as<int32_t>(), sl.counters, sl.d_defer.as<int32_t>(), s_k, memo));
    mark(2, s_k);
    if (keys_stride > 0 || want_match) {
      // keys_stride == 0 with match / routing requested (ids_stride < block_size): every request has 0 blocks, the
      // match is all-zero and routing takes get_load_metrics' least-loaded fallback, as the reference does for a
      // prompt shorter than one block (global_kvcache_mgr.cpp:77-79, instance_mgr.cpp:312-358)
      prep_rows_kernel<<<(m + 127) / 128, 128, 0, s_k>>>(sl.d_n_ids.as<int32_t>(), m, io->ids_stride, keys_stride,
                                                        h->block_size, sl.d_tok_start.as<int64_t>(),
                                                        sl.d_n_tok.as<int32_t>(), sl.d_key_start.as<int64_t>(),
                                                        sl.d_n_blocks.as<int32_t>());
      PIPE_CUDA_TRY(cudaGetLastError());
      chunk_keys = sharded ? h->d_all_keys.as<uint8_t>() + (size_t)c0 * (size_t)keys_stride * 16 : sl.d_keys.as<uint8_t>();
      if (keys_stride > 0) {
        if (io->keys) PIPE_CUDA_TRY(cudaMemsetAsync(chunk_keys, 0, (size_t)m * (size_t)keys_stride * 16, s_k));
        PIPE_CUDA_TRY(xxh3_chain_launch(sl.d_ids.as<int32_t>(), sl.d_tok_start.as<int64_t>(), sl.d_n_tok.as<int32_t>(),
                                        chunk_keys, sl.d_key_start.as<int64_t>(), m, h->block_size, h->xxh,
                                        sl.counters + 8, s_k));
      }
      if (sharded) {
        batch_rows_kernel<<<(m + 127) / 128, 128, 0, s_k>>>(sl.d_n_blocks.as<int32_t>(), m, (int64_t)c0, keys_stride,
                                                           h->d_all_key_start.as<int64_t>(),
                                                           h->d_all_n_blocks.as<int32_t>());
        PIPE_CUDA_TRY(cudaGetLastError());
      } else if (want_match) {
        // probe + first-miss scan + routing in one kernel, between begin_read / end_read so that a publish from
        // another handle of this index waits for it (prefix_index.cuh)
        h->index->begin_read();
        const cudaError_t me = h->index->match_route(sl.d_keys.as<uint8_t>(), sl.d_key_start.as<int64_t>(),
                                                     sl.d_n_blocks.as<int32_t>(), m, h->d_inst,
                                                     sl.d_match.as<MatchOut>(), sl.d_routing.as<RoutingOut>(), s_k);
        h->index->end_read(h->index_read_ev, s_k);
        PIPE_CUDA_TRY(me);
      }
    }
    PIPE_CUDA_TRY(cudaEventRecord(sl.ev[1], s_k));
    // ---- download
    PIPE_CUDA_TRY(cudaStreamWaitEvent(s_out, sl.ev[1], 0));
    PIPE_CUDA_TRY(cudaMemcpyAsync(io->ids + (size_t)c0 * io->ids_stride, sl.d_ids.p,
                                  (size_t)m * (size_t)io->ids_stride * 4, cudaMemcpyDeviceToHost, s_out));
    mark(3, s_out);
    PIPE_CUDA_TRY(cudaMemcpyAsync(io->n_ids + c0, sl.d_n_ids.p, (size_t)m * 4, cudaMemcpyDeviceToHost, s_out));
    PIPE_CUDA_TRY(cudaMemcpyAsync(io->status + c0, sl.d_status.p, (size_t)m * 4, cudaMemcpyDeviceToHost, s_out));
    if (keys_stride > 0 && io->keys)
      PIPE_CUDA_TRY(cudaMemcpyAsync(io->keys + (size_t)c0 * (size_t)keys_stride * 16, chunk_keys,
                                    (size_t)m * (size_t)keys_stride * 16, cudaMemcpyDeviceToHost, s_out));
    if (want_match && !sharded && io->match)
      PIPE_CUDA_TRY(cudaMemcpyAsync(io->match + c0, sl.d_match.p, (size_t)m * sizeof(MatchOut),
                                    cudaMemcpyDeviceToHost, s_out));
    if (want_match && !sharded && io->routing)
      PIPE_CUDA_TRY(cudaMemcpyAsync(io->routing + c0, sl.d_routing.p, (size_t)m * sizeof(RoutingOut),
                                    cudaMemcpyDeviceToHost, s_out));
    PIPE_CUDA_TRY(cudaEventRecord(sl.ev[2], s_out));
    mark(4, s_out);
    sl.busy = true;
    h->last_chunks += 1;
    h->last_launches += 2 + (keys_stride > 0 || want_match ? 1 : 0) + (keys_stride > 0 ? 1 : 0) + (want_match ? 1 : 0);  // encode x2, prep, hash, match+route
    slot = (slot + 1) % n_slots;
    c0 = c1;
  }
#undef PIPE_CUDA_TRY
  if (sharded && rc == XLLM_OK) {
    // every chunk's hash kernel is queued on s_k ahead of this; the round synchronises s_k before it returns
    rc = h->shard->match_route(*h->index, h->index_read_ev, h->d_all_keys.as<uint8_t>(), h->d_all_key_start.as<int64_t>(),
                               h->d_all_n_blocks.as<int32_t>(), n, (int64_t)n * keys_stride, h->d_inst,
                               h->d_all_match.as<MatchOut>(), h->d_all_routing.as<RoutingOut>(), s_k);
    h->last_launches += 5;   // bucket, headers, owner probe, header gather, scan + route (+ 2 NCCL rounds)
    cudaError_t ce = cudaSuccess;
    if (rc == XLLM_OK && io->match)
      ce = cudaMemcpyAsync(io->match, h->d_all_match.p, (size_t)n * sizeof(MatchOut), cudaMemcpyDeviceToHost, s_k);
    if (rc == XLLM_OK && ce == cudaSuccess && io->routing)
      ce = cudaMemcpyAsync(io->routing, h->d_all_routing.p, (size_t)n * sizeof(RoutingOut), cudaMemcpyDeviceToHost, s_k);
    if (ce != cudaSuccess) {
      set_last_error("xllm_ingest_batch (sharded match download): %s", cudaGetErrorString(ce));
      rc = XLLM_ERR_CUDA;
    }
  }
  {
    cudaError_t e = cudaStreamSynchronize(s_out);
    cudaError_t e2 = cudaStreamSynchronize(s_k);
    cudaError_t e3 = cudaStreamSynchronize(s_in);
    if (e == cudaSuccess) e = e2;
    if (e == cudaSuccess) e = e3;
    if (e != cudaSuccess && rc == XLLM_OK) {
      set_last_error("xllm_ingest_batch: %s", cudaGetErrorString(e));
      rc = XLLM_ERR_CUDA;
    }
  }
  if (trace && rc == XLLM_OK) {
    fprintf(stderr, "# chunk m host_enqueue  h2d_begin h2d_end encode_end d2h_ids_end all_end   (ms since batch start)\n");
    for (size_t i = 0; i < rows.size(); ++i) {
      float t[5];
      for (int k = 0; k < 5; ++k) { cudaEventElapsedTime(&t[k], ev_begin, rows[i].e[k]); cudaEventDestroy(rows[i].e[k]); }
      fprintf(stderr, "%3zu %5d %8.3f  %8.3f %8.3f %8.3f %8.3f %8.3f\n", i, rows[i].m, rows[i].host_ms, t[0], t[1], t[2], t[3], t[4]);
    }
    cudaEventDestroy(ev_begin);
  }
  return rc;
}

int xllm_last_batch_stats(xllm_ingest_t h, int32_t* n_chunks, int32_t* n_kernel_launches) {
  if (!h) return XLLM_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(h->mu);
  if (n_chunks) *n_chunks = h->last_chunks;
  if (n_kernel_launches) *n_kernel_launches = h->last_launches;
  return XLLM_OK;
}

int xllm_set_pipeline(xllm_ingest_t h, int32_t chunk_requests, int64_t chunk_bytes) {
  if (!h || chunk_requests <= 0 || chunk_bytes <= 0) return XLLM_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(h->mu);
  h->pipe_chunk_req = chunk_requests;
  h->pipe_chunk_bytes = chunk_bytes;
  return XLLM_OK;
}

}  // extern "C"
