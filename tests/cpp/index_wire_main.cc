// Pins xllm_service_b200/host/index_wire.h against the real nlohmann::json the reference serialises with
// (common/types.h:320-365 CacheLocations::serialize_to_json / parse_from_json; etcd_client.cpp:122-137,174-198).
// The struct below is the reference's CacheLocations restated over the real library: same member types, same calls.
#include <nlohmann/json.hpp>

#include <cstdio>
#include <random>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../xllm_service_b200/host/index_wire.h"

struct RefCacheLocations {
  std::unordered_set<std::string> hbm_instance_set, dram_instance_set, ssd_instance_set;
  nlohmann::json serialize_to_json() const {
    nlohmann::json json_val;
    json_val["hbm_instance_set"] = hbm_instance_set;
    json_val["dram_instance_set"] = dram_instance_set;
    json_val["ssd_instance_set"] = ssd_instance_set;
    return json_val;
  }
  bool parse_from_json(const std::string& json_str) {
    try {
      nlohmann::json json_value = nlohmann::json::parse(json_str);
      for (const auto& item : json_value.at("hbm_instance_set").get<std::vector<std::string>>()) hbm_instance_set.insert(item);
      for (const auto& item : json_value.at("dram_instance_set").get<std::vector<std::string>>()) dram_instance_set.insert(item);
      for (const auto& item : json_value.at("ssd_instance_set").get<std::vector<std::string>>()) ssd_instance_set.insert(item);
    } catch (const std::exception&) {
      return false;
    }
    return true;
  }
};

static int fails = 0;
#define EXPECT(c)                                                   \
  do {                                                              \
    if (!(c)) { ++fails; printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); } \
  } while (0)

static std::unordered_set<std::string> names_of(uint64_t m, const std::vector<std::string>& names) {
  std::unordered_set<std::string> s;
  for (int i = 0; i < 64; ++i)
    if ((m >> i) & 1) s.insert(names[i]);
  return s;
}

int main() {
  std::mt19937_64 rng(7);
  std::vector<std::string> names;
  const char* odd[] = {"10.0.0.1:8000", "instance-\"quoted\"", "back\\slash", "tab\there", "nl\nname", "\x01\x1f ctl",
                       "caf\xc3\xa9", "\xe6\x97\xa5\xe6\x9c\xac", "\xf0\x9f\x99\x82", "", "sp ace", "del\x7f", "/slash"};
  for (int i = 0; i < 64; ++i) names.push_back(i < 13 ? odd[i] : "instance-" + std::to_string(i));
  std::unordered_map<std::string, int> ids;
  for (int i = 0; i < 64; ++i) ids[names[i]] = i;
  auto id_of = [&](const std::string& n) { auto it = ids.find(n); return it == ids.end() ? -1 : it->second; };

  for (int it = 0; it < 4000; ++it) {
    uint64_t m[3];
    for (auto& x : m) {
      x = rng();
      const int keep = (int)(rng() % 5);  // sparse and dense sets, empty ones too
      if (keep == 0) x = 0;
      else if (keep < 3) x &= rng() & rng();
    }
    // (1) our text is exactly nlohmann's compact dump of the same object with the names in ascending-id order
    std::string ours;
    EXPECT(xllm_host::cache_locations_to_json(m[0], m[1], m[2], names, &ours));
    nlohmann::json ordered;
    for (int w = 0; w < 3; ++w) {
      std::vector<std::string> v;
      for (int i = 0; i < 64; ++i)
        if ((m[w] >> i) & 1) v.push_back(names[i]);
      ordered[w == 0 ? "hbm_instance_set" : (w == 1 ? "dram_instance_set" : "ssd_instance_set")] = v;
    }
    EXPECT(ours == ordered.dump());
    // (2) the reference reads our text back to the same three sets
    RefCacheLocations back;
    EXPECT(back.parse_from_json(ours));
    EXPECT(back.hbm_instance_set == names_of(m[0], names) && back.dram_instance_set == names_of(m[1], names) &&
           back.ssd_instance_set == names_of(m[2], names));
    // (3) what the reference writes (unordered_set order, its own dump) parses to the same masks here
    RefCacheLocations ref;
    ref.hbm_instance_set = names_of(m[0], names);
    ref.dram_instance_set = names_of(m[1], names);
    ref.ssd_instance_set = names_of(m[2], names);
    const std::string theirs = ref.serialize_to_json().dump();
    uint64_t g[3] = {1, 2, 3};
    EXPECT(xllm_host::cache_locations_from_json(theirs, id_of, &g[0], &g[1], &g[2]));
    EXPECT(g[0] == m[0] && g[1] == m[1] && g[2] == m[2]);
    // also the pretty-printed and ASCII-escaped forms nlohmann can emit
    EXPECT(xllm_host::cache_locations_from_json(ref.serialize_to_json().dump(2), id_of, &g[0], &g[1], &g[2]));
    EXPECT(g[0] == m[0] && g[1] == m[1] && g[2] == m[2]);
    EXPECT(xllm_host::cache_locations_from_json(ref.serialize_to_json().dump(-1, ' ', true), id_of, &g[0], &g[1], &g[2]));
    EXPECT(g[0] == m[0] && g[1] == m[1] && g[2] == m[2]);
  }
  // (4) both sides reject / accept the same malformed or unusual documents
  const char* docs[] = {
      "{}", "{\"hbm_instance_set\":[]}", "{\"hbm_instance_set\":[],\"dram_instance_set\":[],\"ssd_instance_set\":[]}",
      "{\"hbm_instance_set\":[],\"dram_instance_set\":[],\"ssd_instance_set\":[],\"extra\":{\"a\":[1,2,{\"b\":null}]}}",
      "{\"hbm_instance_set\":[1],\"dram_instance_set\":[],\"ssd_instance_set\":[]}",
      "{\"hbm_instance_set\":\"x\",\"dram_instance_set\":[],\"ssd_instance_set\":[]}",
      "{\"hbm_instance_set\":[\"instance-20\",\"instance-20\"],\"dram_instance_set\":[],\"ssd_instance_set\":[]}",
      " {\n \"ssd_instance_set\" : [ \"instance-21\" ] ,\"dram_instance_set\":[],\"hbm_instance_set\":[\"instance-22\"] } ",
      "{\"hbm_instance_set\":[],\"dram_instance_set\":[],\"ssd_instance_set\":[]} trailing",
      "{\"hbm_instance_set\":[],\"dram_instance_set\":[],\"ssd_instance_set\":[],}",
      "[\"hbm_instance_set\"]", "", "null", "{\"hbm_instance_set\":[\"a\\qb\"],\"dram_instance_set\":[],\"ssd_instance_set\":[]}",
      "{\"hbm_instance_set\":[\"\\ud83d\\ude42\"],\"dram_instance_set\":[],\"ssd_instance_set\":[]}",
      "{\"hbm_instance_set\":[\"\\ud83d\"],\"dram_instance_set\":[],\"ssd_instance_set\":[]}",
      "{\"hbm_instance_set\":[\"raw\ttab\"],\"dram_instance_set\":[],\"ssd_instance_set\":[]}",
      "{\"hbm_instance_set\":[],\"hbm_instance_set\":[\"instance-23\"],\"dram_instance_set\":[],\"ssd_instance_set\":[]}"};
  ids["\xf0\x9f\x99\x82"] = 8;
  for (const char* d : docs) {
    RefCacheLocations ref;
    const bool rok = ref.parse_from_json(d);
    uint64_t g[3] = {0, 0, 0};
    auto any_id = [&](const std::string& n) { auto it = ids.find(n); return it == ids.end() ? -1 : it->second; };
    const bool ok = xllm_host::cache_locations_from_json(d, any_id, &g[0], &g[1], &g[2]);
    if (rok != ok) { ++fails; printf("FAIL accept/reject differs (ref %d, ours %d) on: %s\n", (int)rok, (int)ok, d); }
    if (rok && ok) EXPECT(ref.hbm_instance_set == names_of(g[0], names) && ref.dram_instance_set == names_of(g[1], names) &&
                          ref.ssd_instance_set == names_of(g[2], names));
  }
  // (5) invalid UTF-8 in a name: nlohmann throws, we refuse
  {
    std::vector<std::string> bad = names;
    bad[0] = "broken\xff";
    std::string out;
    EXPECT(!xllm_host::cache_locations_to_json(1, 0, 0, bad, &out));
    nlohmann::json j;
    j["x"] = bad[0];
    bool threw = false;
    try { (void)j.dump(); } catch (const std::exception&) { threw = true; }
    EXPECT(threw);
  }
  // (6) etcd keys: namespace + "XLLM:CACHE:" + 16 raw bytes (NULs and all), cut back at the prefix length
  {
    uint8_t key[16];
    for (int i = 0; i < 16; ++i) key[i] = (uint8_t)(i * 17 % 251);
    key[3] = 0;
    key[9] = 0;
    const std::string ns = "/prod/";
    const std::string k = xllm_host::cache_etcd_key(ns, key);
    EXPECT(k.size() == ns.size() + 11 + 16 && k.compare(0, ns.size() + 11, ns + "XLLM:CACHE:") == 0);
    uint8_t back[16];
    EXPECT(xllm_host::parse_cache_etcd_key(k, ns.size() + 11, back) && memcmp(back, key, 16) == 0);
    EXPECT(!xllm_host::parse_cache_etcd_key(k.substr(0, k.size() - 1), ns.size() + 11, back));
  }
  printf(fails ? "FAILED %d\n" : "OK\n", fails);
  return fails ? 1 : 0;
}
