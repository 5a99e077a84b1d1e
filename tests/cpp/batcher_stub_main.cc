// CPU test of host/ingest_batcher.h's threading logic against a STUB of the three C-ABI entry points it uses
// (no GPU, no library): the stub "tokenises" a prompt into its bytes, sleeps a little like a device call, and
// checks that it is never entered concurrently.  32 threads x many requests: every caller must get its own
// prompt back, requests must really be coalesced, nothing may deadlock.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "xllm_ingest.h"

static std::atomic<int> g_inside{0}, g_overlap{0}, g_calls{0}, g_max_batch{0};

extern "C" {
int xllm_set_memo_policy(xllm_ingest_t, int64_t) { return XLLM_OK; }
int xllm_host_alloc(void** out, size_t bytes) { *out = malloc(bytes ? bytes : 1); return *out ? XLLM_OK : XLLM_ERR_NOMEM; }
void xllm_host_free(void* p) { free(p); }
static int g_vocab = 8000;   // < 65536: the batcher takes the narrow (uint16) id download; 0x20000: the int32 one
int xllm_vocab_size(xllm_ingest_t, int32_t* out) { *out = g_vocab; return XLLM_OK; }
int xllm_ingest_batch(xllm_ingest_t, const xllm_ingest_io* io) {
  if (g_inside.fetch_add(1) != 0) g_overlap.fetch_add(1);
  g_calls.fetch_add(1);
  int m = g_max_batch.load();
  while (io->n_req > m && !g_max_batch.compare_exchange_weak(m, io->n_req)) {}
  for (int r = 0; r < io->n_req; ++r) {
    const int64_t b = io->offsets[r], e = io->offsets[r + 1];
    const int64_t n = e - b;
    io->n_ids[r] = (int32_t)n;
    io->status[r] = n > io->ids_stride ? XLLM_ENC_TRUNCATED : XLLM_OK;
    for (int64_t k = 0; k < n && k < io->ids_stride; ++k) {
      if (io->ids_u16) io->ids_u16[(size_t)r * io->ids_stride + k] = io->text[b + k];
      else io->ids[(size_t)r * io->ids_stride + k] = io->text[b + k];
    }
    if (io->routing) { io->routing[r].ok = 1; io->routing[r].prefill_id = (int32_t)(n % 7); io->routing[r].decode_id = -1; }
    if (io->match) io->match[r].max_block_num = (uint32_t)(n / 128);
  }
  std::this_thread::sleep_for(std::chrono::microseconds(300));
  g_inside.fetch_sub(1);
  return XLLM_OK;
}
}

#include "ingest_batcher.h"

// offline_pct: share of the requests submitted as offline (best effort, deferred behind online ones); 100 = offline
// only (nobody to ride with: every request must fall back to the front door after its defer window)
static bool run_config(int n_threads, int per_thread, int max_batch, int max_wait_us, int offline_pct = 0,
                       int offline_defer_us = 3000) {
  g_calls = 0;
  g_max_batch = 0;
  g_overlap = 0;
  xllm_host::IngestBatcher batcher(nullptr, max_batch, 1 << 16, 256, 128, max_wait_us, true, offline_defer_us);
  if (!batcher.ok()) { printf("alloc failed\n"); return false; }
  std::atomic<int> bad{0};
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; ++t)
    th.emplace_back([&, t] {
      for (int i = 0; i < per_thread; ++i) {
        std::string p = "thread " + std::to_string(t) + " request " + std::to_string(i) + std::string((size_t)((t * 7 + i) % 40), 'x');
        if (i % 97 == 0) p.assign(300, 'y');          // longer than max_tokens: truncated result
        if (i % 131 == 0) p.clear();                  // empty prompt
        xllm_host::IngestResult r;
        batcher.submit(p, &r, /*offline=*/(t * 31 + i * 17) % 100 < offline_pct);
        const size_t keep = p.size() < 256 ? p.size() : 256;
        bool ok = r.status == (p.size() > 256 ? XLLM_ENC_TRUNCATED : XLLM_OK) && r.token_ids.size() == keep &&
                  r.routing.ok == 1 && r.routing.prefill_id == (int32_t)(p.size() % 7);
        for (size_t k = 0; ok && k < keep; ++k) ok = r.token_ids[k] == (uint8_t)p[k];
        if (!ok) bad.fetch_add(1);
      }
    });
  for (auto& t : th) t.join();
  // a prompt larger than the staging buffer is refused, not queued forever
  xllm_host::IngestResult big;
  batcher.submit(std::string((1 << 16) + 1, 'z'), &big);
  const bool big_ok = big.status == XLLM_ERR_CAPACITY;
  const unsigned long long total = (unsigned long long)n_threads * per_thread;
  printf("threads=%d max_batch=%d wait=%dus offline=%d%%: requests=%llu batches=%llu calls=%d largest=%d overlap=%d bad=%d "
         "big_refused=%d piggybacked=%llu\n",
         n_threads, max_batch, max_wait_us, offline_pct, total, (unsigned long long)batcher.batches(), g_calls.load(),
         g_max_batch.load(), g_overlap.load(), bad.load(), (int)big_ok, (unsigned long long)batcher.offline_piggybacked());
  const bool mix_ok = offline_pct == 0 ? batcher.offline_piggybacked() == 0
                                       : (offline_pct == 100 ? true : batcher.offline_piggybacked() > 0);
  return bad.load() == 0 && g_overlap.load() == 0 && batcher.requests() == total && batcher.batches() < total / 3 &&
         g_max_batch.load() <= max_batch && g_max_batch.load() > 1 && big_ok && mix_ok;
}

int main() {
  bool pass = run_config(32, 300, 64, 200);       // everybody fits one batch
  pass = run_config(100, 60, 16, 50) && pass;     // far more threads than a batch holds: room waits, two sets busy
  pass = run_config(48, 100, 8, 0) && pass;       // no wait window at all
  pass = run_config(64, 120, 32, 100, 30) && pass;        // config 5's 70:30 online / offline mix: offline ones ride along
  pass = run_config(32, 40, 16, 50, 100, 1000) && pass;   // offline only: nobody to ride with, no starvation
  g_vocab = 0x20000;                                      // large vocabulary: int32 ids end to end
  pass = run_config(32, 100, 64, 200, 30) && pass;
  printf(pass ? "OK\n" : "FAILED\n");
  return pass ? 0 : 1;
}
