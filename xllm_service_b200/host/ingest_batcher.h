// ingest_batcher.h — host-side micro-batching layer in front of xllm_ingest_batch.
//
// The reference runs the ingest path synchronously, one request per brpc worker thread
// (xllm_service/http_service/service.cpp:396,474 -> Scheduler::schedule, scheduler.cpp:107-153; up to 32
// workers / 128 concurrent requests, global_gflags.cpp:32-36).  A GPU wants thousands of requests
// per launch, so worker threads hand their request to this batcher and block until the batch that
// carries it has gone through the device: the call stays "one request in -> ids + routing out"
// (SURVEY.md §8f rank 2).  Header-only, depends only on the C-ABI.
//
// Policy: the first waiting thread becomes the leader; it collects requests for at most
// `max_wait_us` or until `max_batch` requests / `max_bytes` text bytes are queued, runs ONE
// xllm_ingest_batch over page-locked staging buffers, and wakes the others.
#pragma once
#include <stdint.h>
#include <string.h>

#include <chrono>
#include <condition_variable>
#include <mutex>
#include <string_view>
#include <vector>

#include "xllm_ingest.h"

namespace xllm_host {

struct IngestResult {
  int32_t status = XLLM_ERR_CUDA;     // 0 ok, XLLM_ENC_TRUNCATED, or a negative XLLM_ERR_*
  std::vector<int32_t> token_ids;     // Request::token_ids (request/request.h)
  xllm_routing_out routing{};         // Request::routing as instance ids
  xllm_match_out match{};             // OverlapScores
};

class IngestBatcher {
 public:
  // `h` must outlive the batcher.  max_tokens bounds the ids returned per request.
  IngestBatcher(xllm_ingest_t h, int max_batch, size_t max_bytes, int max_tokens, int block_size, int max_wait_us,
                bool want_routing)
      : h_(h), max_batch_(max_batch), max_bytes_(max_bytes), max_tokens_(max_tokens),
        keys_stride_(max_tokens / block_size), max_wait_us_(max_wait_us), want_routing_(want_routing) {
    ok_ = xllm_host_alloc(reinterpret_cast<void**>(&text_), max_bytes) == XLLM_OK &&
          xllm_host_alloc(reinterpret_cast<void**>(&offsets_), sizeof(int64_t) * (max_batch + 1)) == XLLM_OK &&
          xllm_host_alloc(reinterpret_cast<void**>(&ids_), sizeof(int32_t) * (size_t)max_batch * max_tokens) == XLLM_OK &&
          xllm_host_alloc(reinterpret_cast<void**>(&n_ids_), sizeof(int32_t) * max_batch) == XLLM_OK &&
          xllm_host_alloc(reinterpret_cast<void**>(&status_), sizeof(int32_t) * max_batch) == XLLM_OK &&
          xllm_host_alloc(reinterpret_cast<void**>(&match_), sizeof(xllm_match_out) * max_batch) == XLLM_OK &&
          xllm_host_alloc(reinterpret_cast<void**>(&routing_), sizeof(xllm_routing_out) * max_batch) == XLLM_OK;
  }
  ~IngestBatcher() {
    xllm_host_free(text_); xllm_host_free(offsets_); xllm_host_free(ids_); xllm_host_free(n_ids_);
    xllm_host_free(status_); xllm_host_free(match_); xllm_host_free(routing_);
  }
  bool ok() const { return ok_; }
  uint64_t batches() const { return n_batches_; }
  uint64_t requests() const { return n_requests_; }

  // Blocks until the request has been tokenised (+ matched and routed).  Thread-safe.
  void submit(std::string_view prompt, IngestResult* out) {
    std::unique_lock<std::mutex> lk(mu_);
    if (!ok_ || prompt.size() > max_bytes_) { out->status = XLLM_ERR_CAPACITY; return; }
    // wait for room in the batch that is being assembled
    cv_room_.wait(lk, [&] { return !running_ && (int)pending_.size() < max_batch_ && bytes_ + prompt.size() <= max_bytes_; });
    const uint64_t my_gen = gen_;
    pending_.push_back(Item{prompt, out});
    bytes_ += prompt.size();
    if (pending_.size() == 1) {
      // leader: give followers max_wait_us to join, or go as soon as the batch is full
      const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(max_wait_us_);
      cv_full_.wait_until(lk, deadline, [&] { return (int)pending_.size() >= max_batch_ || bytes_ >= max_bytes_; });
      run_batch(lk);
    } else {
      if ((int)pending_.size() >= max_batch_ || bytes_ >= max_bytes_) cv_full_.notify_one();
      cv_done_.wait(lk, [&] { return gen_ != my_gen; });
    }
  }

 private:
  struct Item {
    std::string_view prompt;
    IngestResult* out;
  };
  void run_batch(std::unique_lock<std::mutex>& lk) {
    std::vector<Item> batch;
    batch.swap(pending_);
    bytes_ = 0;
    running_ = true;
    lk.unlock();
    const int n = (int)batch.size();
    int64_t at = 0;
    for (int i = 0; i < n; ++i) {
      offsets_[i] = at;
      memcpy(text_ + at, batch[i].prompt.data(), batch[i].prompt.size());
      at += (int64_t)batch[i].prompt.size();
    }
    offsets_[n] = at;
    xllm_ingest_io io;
    memset(&io, 0, sizeof(io));
    io.n_req = n;
    io.text = text_;
    io.offsets = offsets_;
    io.ids = ids_;
    io.ids_stride = max_tokens_;
    io.n_ids = n_ids_;
    io.status = status_;
    io.match = want_routing_ ? match_ : nullptr;
    io.routing = want_routing_ ? routing_ : nullptr;
    const int rc = xllm_ingest_batch(h_, &io);
    for (int i = 0; i < n; ++i) {
      IngestResult* o = batch[i].out;
      o->status = rc != XLLM_OK ? rc : status_[i];
      if (rc == XLLM_OK && status_[i] >= 0) {
        const int32_t keep = n_ids_[i] < max_tokens_ ? n_ids_[i] : max_tokens_;
        o->token_ids.assign(ids_ + (size_t)i * max_tokens_, ids_ + (size_t)i * max_tokens_ + keep);
        if (want_routing_) { o->match = match_[i]; o->routing = routing_[i]; }
      }
    }
    lk.lock();
    running_ = false;
    ++gen_;
    ++n_batches_;
    n_requests_ += (uint64_t)n;
    cv_done_.notify_all();
    cv_room_.notify_all();
  }

  xllm_ingest_t h_;
  const int max_batch_;
  const size_t max_bytes_;
  const int max_tokens_;
  const int keys_stride_;
  const int max_wait_us_;
  const bool want_routing_;
  bool ok_ = false;
  uint8_t* text_ = nullptr;
  int64_t* offsets_ = nullptr;
  int32_t* ids_ = nullptr;
  int32_t* n_ids_ = nullptr;
  int32_t* status_ = nullptr;
  xllm_match_out* match_ = nullptr;
  xllm_routing_out* routing_ = nullptr;
  std::mutex mu_;
  std::condition_variable cv_room_, cv_full_, cv_done_;
  std::vector<Item> pending_;
  size_t bytes_ = 0;
  bool running_ = false;
  uint64_t gen_ = 0, n_batches_ = 0, n_requests_ = 0;
};

}  // namespace xllm_host
