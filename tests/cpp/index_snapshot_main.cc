// GPU test driver for host/index_snapshot.h: master table -> etcd pairs -> replica table, then watch-style events.
// Prints "OK <live keys>" or a failure line.  Run by tests/test_gpu_index_snapshot.py.
#include <algorithm>
#include <array>
#include <cstdio>
#include <cstring>
#include <map>
#include <random>
#include <unordered_map>

#include "index_snapshot.h"

using Row = std::array<uint64_t, 5>;
static bool dump(xllm_ingest_t h, std::map<std::pair<uint64_t, uint64_t>, std::array<uint64_t, 3>>* out) {
  int64_t n = 0;
  if (xllm_index_size(h, &n) != XLLM_OK) return false;
  std::vector<uint8_t> keys((size_t)n * 16 + 16);
  std::vector<uint64_t> a((size_t)n + 1), b(a.size()), c(a.size());
  int64_t got = 0;
  if (xllm_index_export(h, n, keys.data(), a.data(), b.data(), c.data(), &got) != XLLM_OK || got != n) return false;
  out->clear();
  for (int64_t i = 0; i < got; ++i) {
    uint64_t k[2];
    memcpy(k, keys.data() + 16 * i, 16);
    (*out)[{k[0], k[1]}] = {a[(size_t)i], b[(size_t)i], c[(size_t)i]};
  }
  return (int64_t)out->size() == got;
}

int main() {
  xllm_ingest_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.block_size = 128;
  cfg.xxh3_seed = 1024;
  cfg.index_capacity = 1 << 16;
  xllm_ingest_t master = nullptr, replica = nullptr;
  if (xllm_ingest_create(&cfg, &master) != XLLM_OK || xllm_ingest_create(&cfg, &replica) != XLLM_OK) {
    printf("create failed: %s\n", xllm_last_error());
    return 1;
  }
  std::vector<std::string> names;
  for (int i = 0; i < 40; ++i) names.push_back(i == 7 ? "odd \"name\"\t\xc3\xa9" : "10.0.0." + std::to_string(i) + ":9000");
  std::mt19937_64 rng(99);
  // master: heartbeat events with stored / offload / removed keys over 40 instances, several publishes
  std::vector<std::array<uint8_t, 16>> pool(20000);
  for (auto& k : pool) for (auto& b : k) b = (uint8_t)rng();
  for (int round = 0; round < 6; ++round) {
    for (int ev = 0; ev < 200; ++ev) {
      std::vector<uint8_t> st, off, rem;
      for (int j = 0; j < 60; ++j) { const auto& k = pool[rng() % pool.size()]; st.insert(st.end(), k.begin(), k.end()); }
      for (int j = 0; j < 15; ++j) { const auto& k = pool[rng() % pool.size()]; off.insert(off.end(), k.begin(), k.end()); }
      for (int j = 0; j < 10; ++j) { const auto& k = pool[rng() % pool.size()]; rem.insert(rem.end(), k.begin(), k.end()); }
      if (xllm_index_apply(master, (int)(rng() % 40), st.data(), st.size() / 16, off.data(), off.size() / 16, rem.data(),
                           rem.size() / 16) != XLLM_OK) { printf("apply failed\n"); return 1; }
    }
    if (xllm_index_publish(master) != XLLM_OK) { printf("publish failed: %s\n", xllm_last_error()); return 1; }
  }
  // master -> etcd pairs
  const std::string ns = "/cluster-a/";
  std::vector<xllm_host::CacheKv> kvs;
  if (xllm_host::snapshot_index(master, ns, names, &kvs) != XLLM_OK) { printf("snapshot failed: %s\n", xllm_last_error()); return 1; }
  // a few pairs the reference would log and skip
  kvs.push_back({ns + "XLLM:CACHE:short", "{}"});
  kvs.push_back({xllm_host::cache_etcd_key(ns, pool[0].data()) + "x", "{\"hbm_instance_set\":[1]}"});
  // replica: its own name table, registered in the order names are first seen (ids differ from the master's!)
  std::unordered_map<std::string, int> rid;
  std::vector<std::string> rnames;
  auto id_of = [&](const std::string& n) {
    auto it = rid.find(n);
    if (it != rid.end()) return it->second;
    if (rnames.size() >= 64) return -1;
    rid[n] = (int)rnames.size();
    rnames.push_back(n);
    return (int)rnames.size() - 1;
  };
  size_t skipped = 0;
  if (xllm_host::apply_etcd_pairs(replica, ns.size() + 11, kvs, id_of, &skipped) != XLLM_OK) { printf("restore failed: %s\n", xllm_last_error()); return 1; }
  if (skipped != 2) { printf("expected 2 skipped pairs, got %zu\n", skipped); return 1; }
  // compare by instance NAME sets
  std::map<std::pair<uint64_t, uint64_t>, std::array<uint64_t, 3>> a, b;
  if (!dump(master, &a) || !dump(replica, &b) || a.size() != b.size()) { printf("dump mismatch %zu %zu\n", a.size(), b.size()); return 1; }
  auto to_names = [](uint64_t m, const std::vector<std::string>& nm) {
    std::vector<std::string> v;
    for (int i = 0; i < 64; ++i) if ((m >> i) & 1) v.push_back(nm[(size_t)i]);
    std::sort(v.begin(), v.end());
    return v;
  };
  for (const auto& kv : a) {
    auto it = b.find(kv.first);
    if (it == b.end()) { printf("key missing on the replica\n"); return 1; }
    for (int w = 0; w < 3; ++w)
      if (to_names(kv.second[w], names) != to_names(it->second[w], rnames)) { printf("sets differ\n"); return 1; }
  }
  // one watch response: a PUT that overwrites, and a key that is deleted AND put — the reference applies the
  // response's PUTs first and its DELETEs after them (global_kvcache_mgr.cpp:162-169), so that key ends up gone
  const auto first = *a.begin();
  uint8_t k0[16];
  memcpy(k0, &first.first.first, 8);
  memcpy(k0 + 8, &first.first.second, 8);
  std::vector<xllm_host::CacheKv> ev = {
      {xllm_host::cache_etcd_key(ns, k0), "{\"dram_instance_set\":[],\"hbm_instance_set\":[\"brand-new:1\"],\"ssd_instance_set\":[]}"},
      {xllm_host::cache_etcd_key(ns, pool[1].data()), ""},
      {xllm_host::cache_etcd_key(ns, pool[1].data()), "{\"dram_instance_set\":[\"brand-new:1\"],\"hbm_instance_set\":[],\"ssd_instance_set\":[]}"}};
  if (xllm_host::apply_etcd_pairs(replica, ns.size() + 11, ev, id_of, &skipped) != XLLM_OK || skipped != 0) { printf("events failed\n"); return 1; }
  uint64_t m3[3];
  int32_t found = 0;
  const int nid = rid["brand-new:1"];
  if (xllm_index_get(replica, k0, m3, &found) != XLLM_OK || !found || m3[0] != (1ull << nid) || m3[1] || m3[2]) { printf("PUT event not applied\n"); return 1; }
  if (xllm_index_get(replica, pool[1].data(), m3, &found) != XLLM_OK || found) { printf("DELETE must win inside one response\n"); return 1; }
  // the next response puts it back
  std::vector<xllm_host::CacheKv> ev2 = {ev[2]};
  if (xllm_host::apply_etcd_pairs(replica, ns.size() + 11, ev2, id_of, &skipped) != XLLM_OK) { printf("events 2 failed\n"); return 1; }
  if (xllm_index_get(replica, pool[1].data(), m3, &found) != XLLM_OK || !found || m3[1] != (1ull << nid) || m3[0] || m3[2]) { printf("PUT after DELETE not applied\n"); return 1; }
  printf("OK %zu\n", a.size());
  xllm_ingest_destroy(master);
  xllm_ingest_destroy(replica);
  return 0;
}
