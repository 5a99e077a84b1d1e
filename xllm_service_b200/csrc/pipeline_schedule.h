// pipeline_schedule.h — chunk sizes of xllm_ingest_batch (pure host arithmetic, unit-tested on CPU).
//
// Chunk sizes ramp up from chunk_req / 8 to chunk_req: the first chunk's upload + kernels is a stage nothing
// overlaps, so it is kept short; the bulk moves in big chunks (full-GPU kernels, few small copies).  The last
// chunk's download is exposed too, but a tiny last chunk would expose its kernels' latency floor (one warp walks
// one prompt: ~1 ms) instead, so when the remainder allows it the batch ends with one quarter-size chunk behind a
// chunk big enough to cover it.
#pragma once
#include <stdint.h>

namespace xllm {

struct ChunkSchedule {
  int64_t chunk_req;  // upper bound per chunk (>= 1)
  int64_t ramp;       // size the next ramp-up chunk may have
  explicit ChunkSchedule(int64_t chunk_requests) : chunk_req(chunk_requests < 1 ? 1 : chunk_requests) {
    ramp = chunk_req / 8 > 64 ? chunk_req / 8 : (chunk_req < 64 ? chunk_req : 64);
  }
  // requests the next chunk should take when `left` (>= 1) are still to go; always in [1, min(left, chunk_req)]
  int64_t next(int64_t left) {
    const int64_t target = ramp < chunk_req ? ramp : chunk_req;
    const int64_t tail = chunk_req / 4 > 0 ? chunk_req / 4 : 1;
    if (ramp < chunk_req) ramp *= 2;
    if (left <= target) return left;
    if (left <= target + tail && left - tail >= tail) return left - tail;  // then exactly one `tail` chunk remains
    return target;
  }
};

// Memo policy (xllm_set_memo_policy): should the launch about to encode n_req requests clear the table whose age
// (requests since its last clear, -1 = never cleared) is *age?  Updates *age for that launch.
inline bool memo_needs_clear(int64_t persist_requests, int64_t* age, int64_t n_req) {
  const bool clear = persist_requests <= 0 || *age < 0 || *age >= persist_requests;
  *age = (clear ? 0 : *age) + n_req;
  return clear;
}

}  // namespace xllm
