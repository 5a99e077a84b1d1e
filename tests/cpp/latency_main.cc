// Service-shaped latency of the ingest path: N caller threads (the reference's brpc workers: 32 threads / 128
// concurrent requests, common/global_gflags.cpp:32-36) each submit one 4 K-token prompt at a time through
// xllm_host::IngestBatcher — exactly what Scheduler::schedule would do per request (scheduler.cpp:107-153) — and wait
// for token ids + routing.  Prints one JSON object: requests/s and the p50 / p90 / p99 / max of the per-request
// submit latency.  Used by bench.py ("service_latency") and tests/test_gpu_host_cpp.py.
// usage: latency_main <tokenizer_dir> <prompts_file> <n_threads> <seconds> <max_batch> <max_wait_us> [offline_pct]
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

#include "ingest_batcher.h"

int main(int argc, char** argv) {
  if (argc < 7) return 2;
  std::vector<std::string> prompts;
  {
    std::ifstream f(argv[2], std::ios::binary);
    for (;;) {   // records: u32 length LE + bytes
      uint32_t n;
      if (!f.read(reinterpret_cast<char*>(&n), 4)) break;
      std::string s(n, '\0');
      f.read(&s[0], n);
      prompts.push_back(std::move(s));
    }
  }
  if (prompts.empty()) return 2;
  const int n_threads = atoi(argv[3]);
  const double seconds = atof(argv[4]);
  const int max_batch = atoi(argv[5]), max_wait_us = atoi(argv[6]);
  const int offline_pct = argc > 7 ? atoi(argv[7]) : 0;
  size_t max_len = 0;
  for (auto& p : prompts) max_len = std::max(max_len, p.size());
  xllm_ingest_config cfg{};
  cfg.tokenizer_path = argv[1];
  cfg.index_capacity = 1 << 16;
  xllm_ingest_t h = nullptr;
  if (xllm_ingest_create(&cfg, &h) != XLLM_OK) {
    fprintf(stderr, "create failed: %s\n", xllm_last_error());
    return 1;
  }
  for (int i = 0; i < 8; ++i) {
    xllm_set_instance(h, i, i % 2 ? 2 : 1, 1);
    xllm_set_load_metrics(h, i, 1, (uint64_t)i, 0.1f * (float)i);
  }
  const int max_tokens = 4096 + 64;
  xllm_host::IngestBatcher batcher(h, max_batch, (size_t)max_batch * (max_len + 64), max_tokens, 128, max_wait_us, true);
  if (!batcher.ok()) return 1;
  std::atomic<bool> stop{false};
  std::atomic<int> bad{0};
  std::vector<std::vector<float>> lat((size_t)n_threads);
  std::vector<std::thread> th;
  const auto t_start = std::chrono::steady_clock::now();
  for (int t = 0; t < n_threads; ++t)
    th.emplace_back([&, t] {
      size_t i = (size_t)t * 7919u;
      xllm_host::IngestResult r;
      while (!stop.load(std::memory_order_relaxed)) {
        const std::string& p = prompts[i++ % prompts.size()];
        const bool offline = (int)((i * 31u + (size_t)t) % 100u) < offline_pct;
        const auto t0 = std::chrono::steady_clock::now();
        batcher.submit(p, &r, offline);
        const auto t1 = std::chrono::steady_clock::now();
        if (r.status != XLLM_OK || r.token_ids.empty() || r.routing.ok != 1) bad.fetch_add(1);
        if (!offline) lat[(size_t)t].push_back(std::chrono::duration<float, std::milli>(t1 - t0).count());
      }
    });
  std::this_thread::sleep_for(std::chrono::duration<double>(seconds));
  stop = true;
  for (auto& t : th) t.join();
  const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
  std::vector<float> all;
  for (auto& v : lat) all.insert(all.end(), v.begin(), v.end());
  std::sort(all.begin(), all.end());
  auto pct = [&](double q) { return all.empty() ? 0.f : all[std::min(all.size() - 1, (size_t)(q * (double)all.size()))]; };
  printf("{\"threads\": %d, \"max_batch\": %d, \"max_wait_us\": %d, \"offline_pct\": %d, \"requests\": %llu, "
         "\"req_per_s\": %.1f, \"batches\": %llu, \"mean_batch\": %.2f, \"p50_ms\": %.3f, \"p90_ms\": %.3f, "
         "\"p99_ms\": %.3f, \"max_ms\": %.3f, \"bad\": %d, \"offline_piggybacked\": %llu}\n",
         n_threads, max_batch, max_wait_us, offline_pct, (unsigned long long)batcher.requests(),
         (double)batcher.requests() / wall, (unsigned long long)batcher.batches(),
         batcher.batches() ? (double)batcher.requests() / (double)batcher.batches() : 0.0, pct(0.5), pct(0.9), pct(0.99),
         all.empty() ? 0.f : all.back(), bad.load(), (unsigned long long)batcher.offline_piggybacked());
  xllm_ingest_destroy(h);
  return bad.load() ? 3 : 0;
}
