// index_snapshot.h — moving the device-resident prefix index to and from its etcd form.
//
// Master side: GlobalKVCacheMgr::upload_kvcache writes every changed key as
//   etcd[ns + "XLLM:CACHE:" + key16] = CacheLocations JSON, or removes it when all three sets are empty
//   (global_kvcache_mgr.cpp:227-247 -> etcd_client.cpp:122-137).
// Replica / restart side: the constructor lists the prefix and inserts every pair it can parse
//   (global_kvcache_mgr.cpp:47-51 -> etcd_client.cpp:174-198), and the watch applies PUT / DELETE events
//   (global_kvcache_mgr.cpp:133-175).
// These helpers do the same against the GPU table through the C-ABI; the etcd client itself stays the reference's.
#pragma once
#include <functional>
#include <string>
#include <vector>

#include "index_wire.h"
#include "xllm_ingest.h"

namespace xllm_host {

struct CacheKv {
  std::string key;    // namespace + "XLLM:CACHE:" + 16 raw bytes
  std::string value;  // CacheLocations JSON; empty = the key is to be deleted
};

// Every live key of the published index as the pair the master keeps in etcd.  names[i] = instance id i.
inline int snapshot_index(xllm_ingest_t h, const std::string& namespace_prefix, const std::vector<std::string>& names,
                          std::vector<CacheKv>* out) {
  int64_t n = 0;
  int rc = xllm_index_size(h, &n);
  if (rc != XLLM_OK) return rc;
  for (;;) {
    std::vector<uint8_t> keys((size_t)(n > 0 ? n : 1) * 16);
    std::vector<uint64_t> hbm((size_t)(n > 0 ? n : 1)), dram(hbm.size()), ssd(hbm.size());
    int64_t got = 0;
    rc = xllm_index_export(h, n, keys.data(), hbm.data(), dram.data(), ssd.data(), &got);
    if (rc == XLLM_ERR_CAPACITY && got > n) { n = got; continue; }  // grew meanwhile
    if (rc != XLLM_OK) return rc;
    out->clear();
    out->reserve((size_t)got);
    for (int64_t i = 0; i < got; ++i) {
      CacheKv kv;
      kv.key = cache_etcd_key(namespace_prefix, keys.data() + 16 * i);
      if (!cache_locations_to_json(hbm[(size_t)i], dram[(size_t)i], ssd[(size_t)i], names, &kv.value))
        return XLLM_ERR_INVALID_ARG;
      out->push_back(std::move(kv));
    }
    return XLLM_OK;
  }
}

// Applies the pairs of ONE etcd listing or ONE watch response to the table and publishes.  A non-empty value is a
// PUT, an empty one a DELETE.  As in GlobalKVCacheMgr::update_kvcache (global_kvcache_mgr.cpp:141-170) the PUTs
// are collected into a map (the last value of a key wins) and applied first, then every DELETE of the response is
// applied — so a key that is both put and deleted in one response ends up deleted, whatever their order.
// Pairs whose key is short or whose JSON the reference's parser would reject are skipped and counted in *n_skipped,
// as the reference logs and continues (etcd_client.cpp:189-192, global_kvcache_mgr.cpp:151-155).
inline int apply_etcd_pairs(xllm_ingest_t h, size_t prefix_len, const std::vector<CacheKv>& kvs,
                            const std::function<int(const std::string&)>& id_of, size_t* n_skipped) {
  std::vector<uint8_t> keys, dels;
  std::vector<uint64_t> hbm, dram, ssd;
  size_t skipped = 0;
  for (const auto& kv : kvs) {
    uint8_t k[16];
    if (!parse_cache_etcd_key(kv.key, prefix_len, k)) { ++skipped; continue; }
    if (kv.value.empty()) {
      dels.insert(dels.end(), k, k + 16);
      continue;
    }
    uint64_t m[3];
    if (!cache_locations_from_json(kv.value, id_of, &m[0], &m[1], &m[2])) { ++skipped; continue; }
    keys.insert(keys.end(), k, k + 16);
    hbm.push_back(m[0]); dram.push_back(m[1]); ssd.push_back(m[2]);
  }
  int rc = XLLM_OK;
  if (!hbm.empty() &&
      (rc = xllm_index_put_bulk(h, (int64_t)hbm.size(), keys.data(), hbm.data(), dram.data(), ssd.data())) != XLLM_OK)
    return rc;
  for (size_t i = 0; i < dels.size(); i += 16)
    if ((rc = xllm_index_erase(h, dels.data() + i)) != XLLM_OK) return rc;
  if (n_skipped) *n_skipped = skipped;
  return xllm_index_publish(h);
}

}  // namespace xllm_host
