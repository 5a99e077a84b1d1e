// pipeline.cu — the whole ingest hot path for a batch of requests behind one C-ABI call:
//   host text -> [H2D] -> tokenize -> block-hash chain -> index probe -> match scan + routing -> [D2H]
// i.e. what Scheduler::schedule does per request between scheduler.cpp:128 and :135
// (Tokenizer::encode, then CacheAwareRouting::select_instances_pair -> GlobalKVCacheMgr::match ->
// cost_function), batched.  The batch is cut into chunks that flow through kPipeSlots independent
// CUDA streams so the PCIe copies of one chunk overlap the kernels of another; the caller's buffers
// should be page-locked (xllm_host_alloc) for the copies to be asynchronous.
#include <string.h>

#include <algorithm>

#include "../../include/xllm_ingest.h"
#include "handle.h"

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
  if ((rc = d_masks.reserve((size_t)n * (size_t)keys_stride * 24 + 64)) != XLLM_OK) return rc;
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
  if (stream) cudaStreamDestroy(stream);
  if (counters) cudaFree(counters);
  stream = nullptr;
  counters = nullptr;
}

}  // namespace xllm

using namespace xllm;

#define XLLM_TRY(expr)              \
  do {                              \
    int _rc = (expr);               \
    if (_rc != XLLM_OK) return _rc; \
  } while (0)

extern "C" {

int xllm_host_alloc(void** out, size_t bytes) {
  if (!out) return XLLM_ERR_INVALID_ARG;
  *out = nullptr;
  cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
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
  if (want_match) {
    std::lock_guard<std::mutex> l2(*h->index_mu);
    XLLM_CUDA_TRY(cudaMemcpyAsync(h->d_inst, h->inst_host.get(), sizeof(InstanceTable), cudaMemcpyHostToDevice,
                                  h->stream));
    XLLM_CUDA_TRY(cudaStreamSynchronize(h->stream));
  }
  // chunking: bounded by request count and by text bytes
  const int chunk_req = h->pipe_chunk_req;
  const int64_t chunk_bytes = h->pipe_chunk_bytes;
  for (int s = 0; s < kPipeSlots; ++s) {
    PipeSlot& sl = h->pipe[s];
    if (!sl.stream) XLLM_CUDA_TRY(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
    if (!sl.counters) XLLM_CUDA_TRY(cudaMalloc(&sl.counters, 64));
  }
  int slot = 0;
  int32_t c0 = 0;
  int rc = XLLM_OK;
  while (c0 < n) {
    int32_t c1 = c0;
    while (c1 < n && c1 - c0 < chunk_req && (c1 == c0 || io->offsets[c1 + 1] - io->offsets[c0] <= chunk_bytes)) ++c1;
    const int m = c1 - c0;
    const int64_t t0 = io->offsets[c0];
    const size_t text_bytes = (size_t)(io->offsets[c1] - t0);
    PipeSlot& sl = h->pipe[slot];
    cudaStream_t s = sl.stream;
    XLLM_CUDA_TRY(cudaStreamSynchronize(s));  // the slot's previous chunk (and its D2H copies) is done
    if ((rc = sl.ensure(text_bytes, m, io->ids_stride, keys_stride)) != XLLM_OK) break;
    if (h->memo_slots && (rc = sl.d_memo.reserve((size_t)h->memo_slots * 32)) != XLLM_OK) break;
    xllm::SpMemo memo;
    memo.table = h->memo_slots ? sl.d_memo.p : nullptr;
    memo.slots = h->memo_slots;
    if (text_bytes)
      XLLM_CUDA_TRY(cudaMemcpyAsync(sl.d_text.p, io->text + t0, text_bytes, cudaMemcpyHostToDevice, s));
    XLLM_CUDA_TRY(cudaMemcpyAsync(sl.d_offsets.p, io->offsets + c0, (size_t)(m + 1) * 8, cudaMemcpyHostToDevice, s));
    XLLM_CUDA_TRY(sp_encode_launch(h->sp_dev->dev(), sl.d_text.as<uint8_t>() - t0, sl.d_offsets.as<int64_t>(), m,
                                   sl.d_ids.as<int32_t>(), io->ids_stride, sl.d_n_ids.as<int32_t>(),
                                   sl.d_status.as<int32_t>(), sl.counters, sl.d_defer.as<int32_t>(), s, memo));
    XLLM_CUDA_TRY(cudaMemcpyAsync(io->ids + (size_t)c0 * io->ids_stride, sl.d_ids.p,
                                  (size_t)m * (size_t)io->ids_stride * 4, cudaMemcpyDeviceToHost, s));
    XLLM_CUDA_TRY(cudaMemcpyAsync(io->n_ids + c0, sl.d_n_ids.p, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
    XLLM_CUDA_TRY(cudaMemcpyAsync(io->status + c0, sl.d_status.p, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
    if (keys_stride > 0) {
      prep_rows_kernel<<<(m + 127) / 128, 128, 0, s>>>(sl.d_n_ids.as<int32_t>(), m, io->ids_stride, keys_stride,
                                                      h->block_size, sl.d_tok_start.as<int64_t>(),
                                                      sl.d_n_tok.as<int32_t>(), sl.d_key_start.as<int64_t>(),
                                                      sl.d_n_blocks.as<int32_t>());
      XLLM_CUDA_TRY(cudaGetLastError());
      if (io->keys) XLLM_CUDA_TRY(cudaMemsetAsync(sl.d_keys.p, 0, (size_t)m * (size_t)keys_stride * 16, s));
      XLLM_CUDA_TRY(xxh3_chain_launch(sl.d_ids.as<int32_t>(), sl.d_tok_start.as<int64_t>(), sl.d_n_tok.as<int32_t>(),
                                      sl.d_keys.as<uint8_t>(), sl.d_key_start.as<int64_t>(), m, h->block_size, h->xxh,
                                      sl.counters + 8, s));
      if (io->keys)
        XLLM_CUDA_TRY(cudaMemcpyAsync(io->keys + (size_t)c0 * (size_t)keys_stride * 16, sl.d_keys.p,
                                      (size_t)m * (size_t)keys_stride * 16, cudaMemcpyDeviceToHost, s));
      if (want_match) {
        XLLM_CUDA_TRY(h->index->probe(sl.d_keys.as<uint8_t>(), (int64_t)m * keys_stride, sl.d_masks.as<uint64_t>(), s));
        XLLM_CUDA_TRY(score_route_launch(sl.d_masks.as<uint64_t>(), sl.d_key_start.as<int64_t>(),
                                         sl.d_n_blocks.as<int32_t>(), m, h->d_inst, sl.d_match.as<MatchOut>(),
                                         sl.d_routing.as<RoutingOut>(), s));
        if (io->match)
          XLLM_CUDA_TRY(cudaMemcpyAsync(io->match + c0, sl.d_match.p, (size_t)m * sizeof(MatchOut),
                                        cudaMemcpyDeviceToHost, s));
        if (io->routing)
          XLLM_CUDA_TRY(cudaMemcpyAsync(io->routing + c0, sl.d_routing.p, (size_t)m * sizeof(RoutingOut),
                                        cudaMemcpyDeviceToHost, s));
      }
    }
    slot = (slot + 1) % kPipeSlots;
    c0 = c1;
  }
  for (int s = 0; s < kPipeSlots; ++s)
    if (h->pipe[s].stream) {
      cudaError_t e = cudaStreamSynchronize(h->pipe[s].stream);
      if (e != cudaSuccess && rc == XLLM_OK) {
        set_last_error("xllm_ingest_batch: %s", cudaGetErrorString(e));
        rc = XLLM_ERR_CUDA;
      }
    }
  return rc;
}

int xllm_set_pipeline(xllm_ingest_t h, int32_t chunk_requests, int64_t chunk_bytes) {
  if (!h || chunk_requests <= 0 || chunk_bytes <= 0) return XLLM_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(h->mu);
  h->pipe_chunk_req = chunk_requests;
  h->pipe_chunk_bytes = chunk_bytes;
  return XLLM_OK;
}

}  // extern "C"
