// CPU unit test of csrc/pipeline_schedule.h: every schedule covers the batch exactly with chunks in [1, chunk_req].
#include <cstdint>
#include <cstdio>
#include <vector>

#include "../../xllm_service_b200/csrc/pipeline_schedule.h"

int main() {
  int fails = 0;
  const long caps[] = {1, 2, 3, 7, 8, 17, 37, 63, 64, 65, 100, 512, 1024, 4096, 8192};
  for (long cap : caps) {
    for (long n = 1; n <= 40000; n += (n < 9000 ? 1 : 97)) {
      xllm::ChunkSchedule s(cap);
      long left = n, chunks = 0;
      std::vector<long> sizes;
      while (left > 0) {
        const long m = (long)s.next(left);
        if (m < 1 || m > left || m > cap) { ++fails; printf("bad chunk %ld (n %ld cap %ld left %ld)\n", m, n, cap, left); break; }
        sizes.push_back(m);
        left -= m;
        if (++chunks > n + 8) { ++fails; printf("no progress n %ld cap %ld\n", n, cap); break; }
      }
      // no more chunks than the ramp (<= 4 steps) plus the bulk needs
      if (chunks > n / cap + 8) { ++fails; printf("too many chunks %ld (n %ld cap %ld)\n", chunks, n, cap); }
      // ramp: sizes never shrink except for the last two chunks
      for (size_t i = 1; i + 2 < sizes.size(); ++i)
        if (sizes[i] < sizes[i - 1]) { ++fails; printf("shrinks early n %ld cap %ld\n", n, cap); break; }
      if (fails > 20) return 1;
    }
  }
  // the documented shape at the defaults
  {
    xllm::ChunkSchedule s(4096);
    long left = 65536;
    std::vector<long> sizes;
    while (left > 0) { sizes.push_back((long)s.next(left)); left -= sizes.back(); }
    const long want_head[] = {512, 1024, 2048, 4096};
    for (int i = 0; i < 4; ++i) if (sizes[(size_t)i] != want_head[i]) { ++fails; printf("head %d = %ld\n", i, sizes[(size_t)i]); }
    if (sizes.back() != 1024) { ++fails; printf("tail = %ld\n", sizes.back()); }
  }
  // memo policy (xllm_set_memo_policy): age-based clearing of a word memo that outlives its launch
  {
    int64_t age = -1;
    if (!xllm::memo_needs_clear(0, &age, 10) || age != 10) { ++fails; printf("policy 0 must clear every launch\n"); }
    if (!xllm::memo_needs_clear(0, &age, 5) || age != 5) { ++fails; printf("policy 0, second launch\n"); }
    age = -1;
    if (!xllm::memo_needs_clear(100, &age, 40) || age != 40) { ++fails; printf("a never-cleared table must be cleared\n"); }
    if (xllm::memo_needs_clear(100, &age, 40) || age != 80) { ++fails; printf("kept below the limit\n"); }
    if (xllm::memo_needs_clear(100, &age, 40) || age != 120) { ++fails; printf("the launch that crosses the limit still keeps it\n"); }
    if (!xllm::memo_needs_clear(100, &age, 7) || age != 7) { ++fails; printf("cleared once it has seen the limit\n"); }
  }
  printf(fails ? "FAILED %d\n" : "OK\n", fails);
  return fails ? 1 : 0;
}
