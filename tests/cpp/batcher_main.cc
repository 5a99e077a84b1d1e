// Drives xllm_host::IngestBatcher from many threads (the brpc-worker pattern) and prints every
// request's token ids + routing; tests/test_gpu_host_cpp.py compares them with the CPU oracle.
// usage: batcher_main <tokenizer_dir> <prompts_file> <n_threads> <max_batch> <max_wait_us>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

#include "ingest_batcher.h"
#include "tokenizers.h"

int main(int argc, char** argv) {
  if (argc < 6) return 2;
  std::vector<std::string> prompts;
  {
    std::ifstream f(argv[2], std::ios::binary);
    // records: u32 length LE + bytes
    for (;;) {
      uint32_t n;
      if (!f.read(reinterpret_cast<char*>(&n), 4)) break;
      std::string s(n, '\0');
      f.read(s.data(), n);
      prompts.push_back(std::move(s));
    }
  }
  xllm_ingest_config cfg{};
  cfg.tokenizer_path = argv[1];
  cfg.index_capacity = 1024;
  xllm_ingest_t h = nullptr;
  if (xllm_ingest_create(&cfg, &h) != XLLM_OK) {
    fprintf(stderr, "create failed: %s\n", xllm_last_error());
    return 1;
  }
  xllm_set_instance(h, 0, 1, 1);
  xllm_set_instance(h, 1, 2, 1);
  xllm_set_load_metrics(h, 0, 1, 3, 0.25f);
  xllm_set_load_metrics(h, 1, 1, 1, 0.5f);
  const int n_threads = atoi(argv[3]);
  xllm_host::IngestBatcher batcher(h, atoi(argv[4]), 1 << 20, 2048, 128, atoi(argv[5]), true);
  if (!batcher.ok()) return 1;
  std::vector<xllm_host::IngestResult> results(prompts.size());
  std::atomic<size_t> next{0};
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; ++t)
    th.emplace_back([&] {
      for (;;) {
        const size_t i = next.fetch_add(1);
        if (i >= prompts.size()) break;
        batcher.submit(prompts[i], &results[i]);
      }
    });
  for (auto& t : th) t.join();
  for (size_t i = 0; i < results.size(); ++i) {
    printf("%d %d %d %zu", results[i].status, results[i].routing.prefill_id, results[i].routing.decode_id,
           results[i].token_ids.size());
    for (int32_t id : results[i].token_ids) printf(" %d", id);
    printf("\n");
  }
  fprintf(stderr, "batches=%llu requests=%llu\n", (unsigned long long)batcher.batches(),
          (unsigned long long)batcher.requests());
  // the legacy ABI from C++ too (what fast_tokenizer.cpp:20-30 does)
  TokenizerHandle th2 = tokenizers_new_from_path(argv[1]);
  TokenizerEncodeResult r;
  tokenizers_encode(th2, prompts[0].data(), prompts[0].size(), 1, &r);
  const bool same = r.len == results[0].token_ids.size() &&
                    std::equal(r.token_ids, r.token_ids + r.len, results[0].token_ids.begin());
  tokenizers_free_encode_results(&r, 1);
  tokenizers_free(th2);
  xllm_ingest_destroy(h);
  return same ? 0 : 3;
}
