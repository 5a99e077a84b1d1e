// Pins xllm_service_b200/host/index_wire.h against the REFERENCE ITSELF (oracle/_ref/libxllm_ref.so = the reference's
// etcd_client.cpp + global_kvcache_mgr.cpp + types.h compiled unmodified over an in-memory etcd):
//  (1) what a reference master writes to etcd in upload_kvcache (etcd_client.cpp:122-137) is read by our key / JSON
//      readers into the same sets the reference holds, and our writer emits the same key and — up to the order of
//      names inside an array (unordered_set iteration in the reference) — the same JSON;
//  (2) pairs produced by OUR writer, PUT into etcd, are loaded by a reference replica's watch
//      (global_kvcache_mgr.cpp:133-175) and by a fresh reference master's get_prefix (etcd_client.cpp:174-198)
//      into exactly the sets we wrote.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../xllm_service_b200/host/index_wire.h"

extern "C" {
void* ref_index_new(uint32_t block_size, uint32_t seed, int is_master, void* share, const char* etcd_namespace);
void ref_index_free(void* h);
long ref_index_size(void* h);
void ref_index_record(void* h, const char* name, const uint8_t* stored, size_t ns, const uint8_t* offload, size_t no,
                      const uint8_t* removed, size_t nr);
int ref_index_upload(void* h);
void ref_etcd_put_raw(void* h, const char* key, size_t key_len, const char* value, size_t value_len);
long ref_etcd_list(void* h, const char* prefix, char* kbuf, size_t kcap, char* vbuf, size_t vcap, int64_t* klen,
                   int64_t* vlen, size_t max_pairs);
int ref_index_get(void* h, const uint8_t* key16, const char* const* names, int n_names, uint64_t* masks3);
}

static int fails = 0;
#define EXPECT(c)                                                   \
  do {                                                              \
    if (!(c)) { ++fails; printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); } \
  } while (0)

static std::string sorted_arrays(const std::string& json) {   // canonical form: sort the strings inside each [...]
  std::string out;
  size_t i = 0;
  while (i < json.size()) {
    if (json[i] != '[') { out += json[i++]; continue; }
    size_t j = i + 1;
    std::vector<std::string> items;
    while (json[j] != ']') {
      if (json[j] == ',') { ++j; continue; }
      size_t s = j++;
      while (json[j] != '"' || json[j - 1] == '\\') {
        if (json[j] == '\\' && json[j + 1] == '\\') ++j;   // skip an escaped backslash pair as a unit
        ++j;
      }
      ++j;
      items.push_back(json.substr(s, j - s));
    }
    std::sort(items.begin(), items.end());
    out += '[';
    for (size_t k = 0; k < items.size(); ++k) out += (k ? "," : "") + items[k];
    out += ']';
    i = j + 1;
  }
  return out;
}

int main() {
  using namespace xllm_host;
  std::mt19937_64 rng(11);
  std::vector<std::string> names;
  const char* odd[] = {"10.0.0.1:8000", "instance-\"quoted\"", "back\\slash", "tab\there", "nl\nname", "caf\xc3\xa9",
                       "\xe6\x97\xa5\xe6\x9c\xac", "sp ace", "/slash"};
  for (int i = 0; i < 40; ++i) names.push_back(i < 9 ? odd[i] : "instance-" + std::to_string(i));
  std::vector<const char*> cnames;
  for (auto& n : names) cnames.push_back(n.c_str());
  std::unordered_map<std::string, int> ids;
  for (int i = 0; i < (int)names.size(); ++i) ids[names[i]] = i;
  auto id_of = [&](const std::string& n) { auto it = ids.find(n); return it == ids.end() ? -1 : it->second; };

  for (const char* ns : {"", "prod/cluster-a"}) {
    const std::string nsp = std::string(ns).empty() ? "" : "/" + std::string(ns) + "/";   // utils.cpp:105-124
    void* master = ref_index_new(128, 1024, 1, nullptr, ns);
    // a random event history over 300 keys
    std::vector<std::vector<uint8_t>> keys(300, std::vector<uint8_t>(16));
    for (auto& k : keys) for (auto& b : k) b = (uint8_t)rng();
    keys[0][3] = 0; keys[1][0] = 0; keys[2][15] = 0;    // embedded NULs must survive (std::string keys)
    for (int round = 0; round < 6; ++round) {
      for (int e = 0; e < 400; ++e) {
        const auto& k = keys[rng() % keys.size()];
        const std::string& n = names[rng() % names.size()];
        int what = (int)(rng() % 10);
        if (what < 6) ref_index_record(master, n.c_str(), k.data(), 1, nullptr, 0, nullptr, 0);
        else if (what < 9) ref_index_record(master, n.c_str(), nullptr, 0, k.data(), 1, nullptr, 0);
        else ref_index_record(master, n.c_str(), nullptr, 0, nullptr, 0, k.data(), 1);
      }
      EXPECT(ref_index_upload(master) == 1);
    }
    // (1) read what the reference wrote
    std::vector<char> kb(1 << 20), vb(1 << 22);
    std::vector<int64_t> kl(4096), vl(4096);
    const std::string prefix = nsp + etcd_cache_prefix();
    long n = ref_etcd_list(master, prefix.c_str(), kb.data(), kb.size(), vb.data(), vb.size(), kl.data(), vl.data(), 4096);
    EXPECT(n > 100 && n == ref_index_size(master));
    struct Row { uint8_t key[16]; uint64_t m[3]; };
    std::vector<Row> rows;
    size_t ko = 0, vo = 0;
    for (long i = 0; i < n; ++i) {
      std::string k(kb.data() + ko, kl[i]), v(vb.data() + vo, vl[i]);
      ko += kl[i]; vo += vl[i];
      Row r;
      EXPECT(parse_cache_etcd_key(k, prefix.size(), r.key));
      EXPECT(cache_etcd_key(nsp, r.key) == k);
      EXPECT(cache_locations_from_json(v, id_of, &r.m[0], &r.m[1], &r.m[2]));
      uint64_t want[3];
      EXPECT(ref_index_get(master, r.key, cnames.data(), (int)cnames.size(), want) == 1);
      EXPECT(want[0] == r.m[0] && want[1] == r.m[1] && want[2] == r.m[2]);
      std::string ours;
      EXPECT(cache_locations_to_json(r.m[0], r.m[1], r.m[2], names, &ours));
      EXPECT(ours.size() == v.size() && sorted_arrays(ours) == sorted_arrays(v));
      rows.push_back(r);
    }
    // (2) our pairs into a second etcd: a replica hears them through its watch, a new master lists them
    void* other = ref_index_new(128, 1024, 1, nullptr, ns);          // owns a fresh store; stays empty itself
    void* replica = ref_index_new(128, 1024, 0, other, ns);
    for (const Row& r : rows) {
      std::string v;
      cache_locations_to_json(r.m[0], r.m[1], r.m[2], names, &v);
      std::string k = cache_etcd_key(nsp, r.key);
      ref_etcd_put_raw(replica, k.data(), k.size(), v.data(), v.size());
    }
    void* late_master = ref_index_new(128, 1024, 1, other, ns);      // constructor get_prefix (:47-51)
    EXPECT(ref_index_size(replica) == (long)rows.size());
    EXPECT(ref_index_size(late_master) == (long)rows.size());
    for (const Row& r : rows) {
      uint64_t a[3], b[3];
      EXPECT(ref_index_get(replica, r.key, cnames.data(), (int)cnames.size(), a) == 1);
      EXPECT(ref_index_get(late_master, r.key, cnames.data(), (int)cnames.size(), b) == 1);
      EXPECT(a[0] == r.m[0] && a[1] == r.m[1] && a[2] == r.m[2]);
      EXPECT(b[0] == r.m[0] && b[1] == r.m[1] && b[2] == r.m[2]);
    }
    ref_index_free(late_master);
    ref_index_free(replica);
    ref_index_free(other);
    ref_index_free(master);
    printf("namespace '%s': %ld pairs checked both ways\n", ns, n);
  }
  printf(fails ? "FAILED %d\n" : "OK\n", fails);
  return fails ? 1 : 0;
}
