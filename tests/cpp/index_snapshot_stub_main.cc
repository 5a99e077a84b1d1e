// CPU test of host/index_snapshot.h against a STUB of the index entry points of the C-ABI (an in-memory map with
// staged writes): the listing / watch-response semantics of GlobalKVCacheMgr (global_kvcache_mgr.cpp:47-51,133-175)
// — PUTs of one response first (last value wins), then its DELETEs; unparsable pairs skipped — and the snapshot's
// key / JSON form, without a GPU.
#include <array>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "xllm_ingest.h"

using Key = std::array<uint8_t, 16>;
struct Masks { uint64_t m[3]; };
static std::map<Key, Masks> g_live;
static std::vector<std::pair<Key, std::pair<bool, Masks>>> g_staged;  // (key, (is_put, masks)) in arrival order
static int g_publishes = 0;

extern "C" {
int xllm_index_put_bulk(xllm_ingest_t, int64_t n, const uint8_t* keys, const uint64_t* a, const uint64_t* b, const uint64_t* c) {
  for (int64_t i = 0; i < n; ++i) {
    Key k;
    memcpy(k.data(), keys + 16 * i, 16);
    g_staged.push_back({k, {true, Masks{{a[i], b[i], c[i]}}}});
  }
  return XLLM_OK;
}
int xllm_index_erase(xllm_ingest_t, const uint8_t* key16) {
  Key k;
  memcpy(k.data(), key16, 16);
  g_staged.push_back({k, {false, Masks{{0, 0, 0}}}});
  return XLLM_OK;
}
int xllm_index_publish(xllm_ingest_t) {
  for (const auto& op : g_staged) {
    if (op.second.first) g_live[op.first] = op.second.second;
    else g_live.erase(op.first);
  }
  g_staged.clear();
  ++g_publishes;
  return XLLM_OK;
}
int xllm_index_size(xllm_ingest_t, int64_t* n) { *n = (int64_t)g_live.size(); return XLLM_OK; }
int xllm_index_export(xllm_ingest_t, int64_t cap, uint8_t* keys, uint64_t* a, uint64_t* b, uint64_t* c, int64_t* n) {
  *n = (int64_t)g_live.size();
  int64_t i = 0;
  for (const auto& kv : g_live) {
    if (i >= cap) break;
    memcpy(keys + 16 * i, kv.first.data(), 16);
    a[i] = kv.second.m[0]; b[i] = kv.second.m[1]; c[i] = kv.second.m[2];
    ++i;
  }
  return *n > cap ? XLLM_ERR_CAPACITY : XLLM_OK;
}
}

#include "index_snapshot.h"

static int fails = 0;
#define EXPECT(c) do { if (!(c)) { ++fails; printf("FAIL line %d: %s\n", __LINE__, #c); } } while (0)

int main() {
  const std::string ns = "/ns/";
  const size_t plen = ns.size() + 11;
  std::unordered_map<std::string, int> ids;
  std::vector<std::string> names;
  auto id_of = [&](const std::string& n) {
    auto it = ids.find(n);
    if (it != ids.end()) return it->second;
    if (names.size() >= 64) return -1;
    ids[n] = (int)names.size();
    names.push_back(n);
    return (int)names.size() - 1;
  };
  auto key = [](int i) { Key k{}; k[0] = (uint8_t)i; k[5] = 0; k[15] = (uint8_t)(i * 3); return k; };
  auto etcd = [&](int i) { return xllm_host::cache_etcd_key(ns, key(i).data()); };
  // start-up listing: three good pairs, one bad JSON, one short key, a duplicate key (last value wins)
  std::vector<xllm_host::CacheKv> listing = {
      {etcd(1), "{\"hbm_instance_set\":[\"a\",\"b\"],\"dram_instance_set\":[],\"ssd_instance_set\":[]}"},
      {etcd(2), "{\"dram_instance_set\":[\"b\"],\"ssd_instance_set\":[\"c\"],\"hbm_instance_set\":[]}"},
      {etcd(3), "{\"hbm_instance_set\":[\"c\"],\"dram_instance_set\":[],\"ssd_instance_set\":[]}"},
      {etcd(4), "{\"hbm_instance_set\":[\"c\"]}"},
      {ns + "XLLM:CACHE:tooshort", "{\"hbm_instance_set\":[],\"dram_instance_set\":[],\"ssd_instance_set\":[]}"},
      {etcd(1), "{\"hbm_instance_set\":[\"a\"],\"dram_instance_set\":[\"c\"],\"ssd_instance_set\":[]}"}};
  size_t skipped = 99;
  EXPECT(xllm_host::apply_etcd_pairs(nullptr, plen, listing, id_of, &skipped) == XLLM_OK);
  EXPECT(skipped == 2 && g_publishes == 1 && g_live.size() == 3);
  EXPECT(g_live[key(1)].m[0] == (1ull << ids["a"]) && g_live[key(1)].m[1] == (1ull << ids["c"]) && g_live[key(1)].m[2] == 0);
  EXPECT(g_live[key(2)].m[1] == (1ull << ids["b"]) && g_live[key(2)].m[2] == (1ull << ids["c"]));
  // one watch response: delete 2, put 2 again, put 5, delete 3 -> puts first, deletes after: 2 and 3 gone, 5 there
  std::vector<xllm_host::CacheKv> resp = {
      {etcd(2), ""},
      {etcd(2), "{\"hbm_instance_set\":[\"d\"],\"dram_instance_set\":[],\"ssd_instance_set\":[]}"},
      {etcd(5), "{\"hbm_instance_set\":[],\"dram_instance_set\":[],\"ssd_instance_set\":[\"a\"]}"},
      {etcd(3), ""}};
  EXPECT(xllm_host::apply_etcd_pairs(nullptr, plen, resp, id_of, &skipped) == XLLM_OK && skipped == 0);
  EXPECT(g_live.size() == 2 && g_live.count(key(1)) && g_live.count(key(5)) && !g_live.count(key(2)) && !g_live.count(key(3)));
  // snapshot: the reference's key and JSON form, readable back into an identical table
  std::vector<xllm_host::CacheKv> snap;
  EXPECT(xllm_host::snapshot_index(nullptr, ns, names, &snap) == XLLM_OK && snap.size() == 2);
  for (const auto& kv : snap) {
    EXPECT(kv.key.compare(0, plen, ns + "XLLM:CACHE:") == 0 && kv.key.size() == plen + 16);
    EXPECT(kv.value.rfind("{\"dram_instance_set\":[", 0) == 0);  // nlohmann's sorted key order, compact
  }
  const auto before = g_live;
  g_live.clear();
  EXPECT(xllm_host::apply_etcd_pairs(nullptr, plen, snap, id_of, &skipped) == XLLM_OK && skipped == 0);
  EXPECT(g_live.size() == before.size());
  for (const auto& kv : before) EXPECT(g_live.count(kv.first) && memcmp(g_live[kv.first].m, kv.second.m, 24) == 0);
  printf(fails ? "FAILED %d\n" : "OK\n", fails);
  return fails ? 1 : 0;
}
