// shard_exchange.cu — see shard_exchange.cuh.
#include "shard_exchange.cuh"

#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <mutex>
#include <string>

#include "common.cuh"

namespace xllm {

namespace {

// ---------------------------------------------------------------------------------------- NCCL through dlopen
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};

NcclApi* nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // a libnccl.so.2 that is already in the process (e.g. the one bundled with torch) is reused by soname
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (!api.lib) {
      api.error = std::string("dlopen(libnccl.so.2) failed: ") + dlerror();
      return;
    }
    auto sym = [&](const char* name) {
      void* p = dlsym(api.lib, name);
      if (!p && api.error.empty()) api.error = std::string("libnccl: missing symbol ") + name;
      return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  });
  return &api;
}

#define XLLM_NCCL_TRY(expr)                                                                               \
  do {                                                                                                    \
    ncclResult_t _r = (expr);                                                                             \
    if (_r != ncclSuccess) {                                                                              \
      ::xllm::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, nccl()->GetErrorString(_r));   \
      return XLLM_ERR_CUDA;                                                                               \
    }                                                                                                     \
  } while (0)

// ---------------------------------------------------------------------------------------- origin-side kernels
constexpr int kMaxWorld = 32;

// Eight request rows per 256-thread block (warp = row, lane = block of the row, waves of 32 blocks).  Per wave the
// block builds a shared histogram of owners, reserves one contiguous range per owner in the outgoing messages with
// a single global atomic each, and writes its tuples there; pos[key index] remembers peer * cap + slot for the way
// back.  A slot beyond `cap` is not written (the round is repeated with a larger capacity): pos = ~0.
__global__ void __launch_bounds__(256) bucket_by_owner_kernel(const uint64_t* __restrict__ keys,
                                                              const int64_t* __restrict__ key_start,
                                                              const int32_t* __restrict__ n_blocks, int n_req,
                                                              int log2_world, int world, int rank, uint32_t cap,
                                                              size_t msg_bytes, uint8_t* __restrict__ send,
                                                              uint32_t* __restrict__ cursors,
                                                              uint32_t* __restrict__ pos) {
  __shared__ uint32_t hist[kMaxWorld], base[kMaxWorld];
  __shared__ int max_nb;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = blockIdx.x * 8 + warp;
  const int nb = r < n_req ? n_blocks[r] : 0;
  const int64_t k0 = r < n_req ? key_start[r] : 0;
  if (threadIdx.x == 0) max_nb = 0;
  __syncthreads();
  if (lane == 0) atomicMax(&max_nb, nb);
  __syncthreads();
  const int waves = (max_nb + 31) >> 5;
  for (int w = 0; w < waves; ++w) {
    if (threadIdx.x < kMaxWorld) hist[threadIdx.x] = 0;
    __syncthreads();
    const int i = w * 32 + lane;
    const bool active = i < nb;
    uint64_t lo = 0, hi = 0;
    int owner = 0;
    uint32_t local = 0;
    if (active) {
      const ulonglong2 k = *reinterpret_cast<const ulonglong2*>(keys + 2 * (k0 + i));
      lo = k.x; hi = k.y;
      owner = log2_world == 0 ? 0 : (int)(lo >> (64 - log2_world));
      local = atomicAdd(&hist[owner], 1u);
    }
    __syncthreads();
    if ((int)threadIdx.x < world && hist[threadIdx.x]) base[threadIdx.x] = atomicAdd(&cursors[threadIdx.x], hist[threadIdx.x]);
    __syncthreads();
    if (active) {
      const uint32_t slot = base[owner] + local;
      if (slot < cap) {
        uint64_t* t = reinterpret_cast<uint64_t*>(send + (size_t)owner * msg_bytes + 16 + (size_t)slot * 24);
        t[0] = lo; t[1] = hi;
        t[2] = (uint64_t)(uint32_t)r | ((uint64_t)(uint16_t)i << 32) | ((uint64_t)(uint16_t)rank << 48);
        pos[k0 + i] = (uint32_t)owner * cap + slot;
      } else {
        pos[k0 + i] = 0xFFFFFFFFu;
      }
    }
    __syncthreads();
  }
}

__global__ void write_headers_kernel(const uint32_t* __restrict__ cursors, int world, uint32_t cap, size_t msg_bytes,
                                     uint8_t* __restrict__ send) {
  const int p = threadIdx.x;
  if (p >= world) return;
  uint32_t mx = 0, total = 0;
  for (int q = 0; q < world; ++q) { mx = cursors[q] > mx ? cursors[q] : mx; total += cursors[q]; }
  ShardHeader h;
  h.count = cursors[p] < cap ? cursors[p] : cap;
  h.max_bucket = mx;
  h.total_keys = total;
  h.pad = 0;
  *reinterpret_cast<ShardHeader*>(send + (size_t)p * msg_bytes) = h;
}

__global__ void gather_headers_kernel(const uint8_t* __restrict__ recv, int world, size_t msg_bytes,
                                      ShardHeader* __restrict__ out) {
  const int p = threadIdx.x;
  if (p < world) out[p] = *reinterpret_cast<const ShardHeader*>(recv + (size_t)p * msg_bytes);
}

}  // namespace

// ---------------------------------------------------------------------------------------- host side
int ShardExchange::unique_id(void* out128) {
  NcclApi* n = nccl();
  if (!n->error.empty()) {
    set_last_error("%s", n->error.c_str());
    return XLLM_ERR_UNSUPPORTED;
  }
  ncclUniqueId id;
  XLLM_NCCL_TRY(n->GetUniqueId(&id));
  static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
  memcpy(out128, &id, sizeof(id));
  return XLLM_OK;
}

int ShardExchange::init(int world, int rank, const void* unique_id128, int device, int64_t bucket_capacity) {
  if (world < 2 || world > kMaxWorld || (world & (world - 1)) != 0 || rank < 0 || rank >= world || !unique_id128 ||
      bucket_capacity <= 0 || bucket_capacity > 0x7FFFFFF0ll / 24) {
    set_last_error("sharded index: world must be a power of two in [2,%d], 0 <= rank < world, a unique id given", kMaxWorld);
    return XLLM_ERR_INVALID_ARG;
  }
  NcclApi* n = nccl();
  if (!n->error.empty()) {
    set_last_error("%s", n->error.c_str());
    return XLLM_ERR_UNSUPPORTED;
  }
  world_ = world;
  rank_ = rank;
  device_ = device;
  log2_ = 0;
  while ((1 << log2_) < world) ++log2_;
  XLLM_CUDA_TRY(cudaSetDevice(device));
  ncclUniqueId id;
  memcpy(&id, unique_id128, sizeof(id));
  ncclComm_t comm = nullptr;
  XLLM_NCCL_TRY(n->CommInitRank(&comm, world, id, rank));
  comm_ = comm;
  XLLM_CUDA_TRY(cudaMalloc(&d_cursors_, (kMaxWorld + 4) * sizeof(uint32_t)));
  XLLM_CUDA_TRY(cudaMalloc(&d_headers_, kMaxWorld * sizeof(ShardHeader)));
  XLLM_CUDA_TRY(cudaHostAlloc(&h_headers_, kMaxWorld * sizeof(ShardHeader), cudaHostAllocDefault));
  for (auto& e : ev_) XLLM_CUDA_TRY(cudaEventCreate(&e));
  cap_ = bucket_capacity;
  return XLLM_OK;
}

ShardExchange::~ShardExchange() {
  if (comm_) nccl()->CommDestroy(static_cast<ncclComm_t>(comm_));
  if (d_send_) cudaFree(d_send_);
  if (d_recv_) cudaFree(d_recv_);
  if (d_back_send_) cudaFree(d_back_send_);
  if (d_back_recv_) cudaFree(d_back_recv_);
  if (d_cursors_) cudaFree(d_cursors_);
  if (d_headers_) cudaFree(d_headers_);
  if (d_pos_) cudaFree(d_pos_);
  if (h_headers_) cudaFreeHost(h_headers_);
  for (auto& e : ev_)
    if (e) cudaEventDestroy(e);
}

int ShardExchange::ensure_buffers(int64_t cap, int64_t n_keys_bound) {
  if (cap > buf_cap_) {
    for (uint8_t** b : {&d_send_, &d_recv_, &d_back_send_, &d_back_recv_}) {
      if (*b) cudaFree(*b);
      *b = nullptr;
    }
    buf_cap_ = 0;
    const size_t msg = 16 + (size_t)cap * 24, back = (size_t)cap * 24;
    XLLM_CUDA_TRY(cudaMalloc(&d_send_, (size_t)world_ * msg));
    XLLM_CUDA_TRY(cudaMalloc(&d_recv_, (size_t)world_ * msg));
    XLLM_CUDA_TRY(cudaMalloc(&d_back_send_, (size_t)world_ * back));
    XLLM_CUDA_TRY(cudaMalloc(&d_back_recv_, (size_t)world_ * back));
    buf_cap_ = cap;
  }
  if (n_keys_bound > pos_cap_) {
    if (d_pos_) cudaFree(d_pos_);
    d_pos_ = nullptr;
    pos_cap_ = 0;
    XLLM_CUDA_TRY(cudaMalloc(&d_pos_, (size_t)n_keys_bound * 4 + 64));
    pos_cap_ = n_keys_bound;
  }
  return XLLM_OK;
}

// one all-to-all of equal-size messages: message p of `send` goes to rank p, message q of `recv` comes from rank q
int ShardExchange::exchange(const uint8_t* send, uint8_t* recv, size_t bytes, cudaStream_t stream) {
  NcclApi* n = nccl();
  ncclComm_t comm = static_cast<ncclComm_t>(comm_);
  XLLM_NCCL_TRY(n->GroupStart());
  for (int p = 0; p < world_; ++p) {
    if (p == rank_) continue;
    XLLM_NCCL_TRY(n->Send(send + (size_t)p * bytes, bytes, ncclUint8, p, comm, stream));
    XLLM_NCCL_TRY(n->Recv(recv + (size_t)p * bytes, bytes, ncclUint8, p, comm, stream));
  }
  XLLM_NCCL_TRY(n->GroupEnd());
  XLLM_CUDA_TRY(cudaMemcpyAsync(recv + (size_t)rank_ * bytes, send + (size_t)rank_ * bytes, bytes,
                                cudaMemcpyDeviceToDevice, stream));
  return XLLM_OK;
}

int ShardExchange::round(PrefixIndex& index, cudaEvent_t index_read_ev, const uint8_t* d_keys,
                         const int64_t* d_key_start, const int32_t* d_n_blocks, int n_req,
                         const InstanceTable* d_instances, MatchOut* d_match, RoutingOut* d_routing,
                         cudaStream_t stream, uint32_t* need_cap) {
  const uint32_t cap = (uint32_t)cap_;
  const size_t msg = 16 + (size_t)cap * 24, back = (size_t)cap * 24;
  XLLM_CUDA_TRY(cudaEventRecord(ev_[0], stream));
  XLLM_CUDA_TRY(cudaMemsetAsync(d_cursors_, 0, (kMaxWorld + 4) * sizeof(uint32_t), stream));
  if (n_req > 0) {
    bucket_by_owner_kernel<<<(n_req + 7) / 8, 256, 0, stream>>>(reinterpret_cast<const uint64_t*>(d_keys), d_key_start,
                                                               d_n_blocks, n_req, log2_, world_, rank_, cap, msg,
                                                               d_send_, d_cursors_, d_pos_);
    XLLM_CUDA_TRY(cudaGetLastError());
  }
  write_headers_kernel<<<1, kMaxWorld, 0, stream>>>(d_cursors_, world_, cap, msg, d_send_);
  XLLM_CUDA_TRY(cudaGetLastError());
  XLLM_CUDA_TRY(cudaEventRecord(ev_[1], stream));
  XLLM_TRY_RC(exchange(d_send_, d_recv_, msg, stream));
  XLLM_CUDA_TRY(cudaEventRecord(ev_[2], stream));
  // owner side: this rank's slice of the table answers every tuple it received
  index.begin_read();
  const cudaError_t pe = index.probe_messages(d_recv_, msg, world_, cap, reinterpret_cast<uint64_t*>(d_back_send_), stream);
  gather_headers_kernel<<<1, kMaxWorld, 0, stream>>>(d_recv_, world_, msg, d_headers_);
  index.end_read(index_read_ev, stream);
  XLLM_CUDA_TRY(pe);
  XLLM_CUDA_TRY(cudaEventRecord(ev_[3], stream));
  XLLM_TRY_RC(exchange(d_back_send_, d_back_recv_, back, stream));
  XLLM_CUDA_TRY(cudaEventRecord(ev_[4], stream));
  XLLM_CUDA_TRY(score_route_launch(reinterpret_cast<const uint64_t*>(d_back_recv_), d_key_start, d_n_blocks, n_req,
                                   d_instances, d_match, d_routing, stream, d_pos_));
  XLLM_CUDA_TRY(cudaEventRecord(ev_[5], stream));
  // headers of the messages received (every rank's largest bucket): one small read-back decides about a repeat
  XLLM_CUDA_TRY(cudaMemcpyAsync(h_headers_, d_headers_, (size_t)world_ * sizeof(ShardHeader), cudaMemcpyDeviceToHost,
                                stream));
  XLLM_CUDA_TRY(cudaStreamSynchronize(stream));
  uint32_t need = 0;
  for (int p = 0; p < world_; ++p) need = h_headers_[p].max_bucket > need ? h_headers_[p].max_bucket : need;
  *need_cap = need;
  float* t = &times_.bucket_ms;
  for (int k = 0; k < 5; ++k) cudaEventElapsedTime(&t[k], ev_[k], ev_[k + 1]);
  return XLLM_OK;
}

int ShardExchange::match_route(PrefixIndex& index, cudaEvent_t index_read_ev, const uint8_t* d_keys,
                               const int64_t* d_key_start, const int32_t* d_n_blocks, int n_req, int64_t n_keys_bound,
                               const InstanceTable* d_instances, MatchOut* d_match, RoutingOut* d_routing,
                               cudaStream_t stream) {
  if (!comm_) {
    set_last_error("sharded index: communicator not initialised");
    return XLLM_ERR_UNSUPPORTED;
  }
  std::lock_guard<std::mutex> lock(mu_);
  XLLM_CUDA_TRY(cudaSetDevice(device_));
  XLLM_TRY_RC(ensure_buffers(cap_, n_keys_bound > 0 ? n_keys_bound : 1));
  uint32_t need = 0;
  XLLM_TRY_RC(round(index, index_read_ev, d_keys, d_key_start, d_n_blocks, n_req, d_instances, d_match, d_routing,
                    stream, &need));
  // A bucket overflowed somewhere.  Every rank received every rank's header, so every rank computes the same `need`
  // and repeats the round with the same capacity — no extra collective to agree on it.
  while ((int64_t)need > cap_) {
    ++overflow_rounds_;
    cap_ = (int64_t)need + (int64_t)need / 8 + 1024;
    XLLM_TRY_RC(ensure_buffers(cap_, n_keys_bound > 0 ? n_keys_bound : 1));
    XLLM_TRY_RC(round(index, index_read_ev, d_keys, d_key_start, d_n_blocks, n_req, d_instances, d_match, d_routing,
                      stream, &need));
  }
  return XLLM_OK;
}

}  // namespace xllm
