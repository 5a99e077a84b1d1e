// reference_adaptors.h — the reference-side bindings a maintainer drops into xllm-service to put the
// B200 path behind its existing seams.  Compiles inside the reference tree (it includes the
// reference's own headers); nothing here is needed by this repo's tests.  See INTEGRATION.md.
//
//   GpuTokenizer          : a 4th Tokenizer backend (xllm_service/tokenizer/tokenizer.h:28-46), selected
//                           in TokenizerFactory::create_tokenizer (tokenizer_factory.cpp:9-32)
//   GpuGlobalKVCacheIndex : the three calls Scheduler / CacheAwareRouting make on GlobalKVCacheMgr
//                           (global_kvcache_mgr.h:39-45) + name <-> instance-id bookkeeping
//   GpuCacheAwareRouting  : LoadBalancePolicy (loadbalance_policy.h:24-35) that fills Request::routing
#pragma once
#include <memory>
#include <mutex>
#include <optional>
#include <string>
#include <unordered_map>
#include <vector>

#include "common/slice.h"
#include "common/types.h"
#include "request/request.h"
#include "scheduler/loadbalance_policy/loadbalance_policy.h"
#include "tokenizer/tokenizer.h"
#include "tokenizers.h"      // this repo: include/tokenizers.h
#include "index_snapshot.h"
#include "xllm_ingest.h"     // this repo: include/xllm_ingest.h
#include "xllm_rpc_service.pb.h"

namespace xllm_service {

// ---------------------------------------------------------------------------- tokenizer
class GpuTokenizer final : public Tokenizer {
 public:
  GpuTokenizer(const std::string& tokenizer_dir, int device, int32_t block_size, uint32_t seed)
      : dir_(tokenizer_dir) {
    xllm_ingest_config cfg{};
    cfg.tokenizer_path = dir_.c_str();
    cfg.block_size = block_size;
    cfg.xxh3_seed = seed;
    cfg.device = device;
    CHECK_EQ(xllm_ingest_create(&cfg, &h_), XLLM_OK) << xllm_last_error();
    legacy_ = tokenizers_new_from_path(dir_.c_str());  // decode / vocabulary queries
  }
  explicit GpuTokenizer(xllm_ingest_t cloned, const std::string& dir) : h_(cloned), dir_(dir) {
    legacy_ = tokenizers_new_from_path(dir_.c_str());
  }
  ~GpuTokenizer() override {
    tokenizers_free(legacy_);
    xllm_ingest_destroy(h_);
  }

  // Appends, like the SentencePiece / tiktoken backends (sentencepiece_tokenizer.cpp:122-126); false on
  // failure so Scheduler::schedule fails exactly as today (scheduler.cpp:129-132).
  bool encode(const std::string_view& text, std::vector<int32_t>* ids) const override {
    if (text.empty()) return true;
    const int64_t off[2] = {0, (int64_t)text.size()};
    std::vector<int32_t> buf(text.size() + 8);
    int32_t n = 0, st = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
      if (xllm_encode_batch(h_, 1, reinterpret_cast<const uint8_t*>(text.data()), off, buf.data(),
                            (int64_t)buf.size(), &n, &st) != XLLM_OK)
        return false;
      if (st != XLLM_ENC_TRUNCATED) break;
      buf.resize((size_t)n);
    }
    if (st != XLLM_OK) return false;
    ids->insert(ids->end(), buf.begin(), buf.begin() + n);
    return true;
  }
  std::string decode(const Slice<int32_t>& ids, bool skip_special_tokens) const override {
    const char* data = nullptr;
    size_t len = 0;
    tokenizers_decode(legacy_, reinterpret_cast<const uint32_t*>(ids.data()), ids.size(), skip_special_tokens, &data,
                      &len);
    return {data, len};
  }
  std::optional<int32_t> token_to_id(const std::string_view& token) const override {
    int32_t id = -1;
    tokenizers_token_to_id(legacy_, token.data(), token.size(), &id);
    return id == -1 ? std::nullopt : std::optional<int32_t>(id);
  }
  std::string id_to_token(int32_t id) const override {
    const char* data = nullptr;
    size_t len = 0;
    tokenizers_id_to_token(legacy_, (uint32_t)id, &data, &len);
    return {data, len};
  }
  size_t vocab_size() const override {
    size_t n = 0;
    tokenizers_get_vocab_size(legacy_, &n);
    return n;
  }
  // Cheap: shares the device tables (the reference's clones reload the model from disk).
  std::unique_ptr<Tokenizer> clone() const override {
    xllm_ingest_t c = nullptr;
    CHECK_EQ(xllm_ingest_clone(h_, &c), XLLM_OK) << xllm_last_error();
    return std::make_unique<GpuTokenizer>(c, dir_);
  }
  xllm_ingest_t handle() const { return h_; }

 private:
  xllm_ingest_t h_ = nullptr;
  TokenizerHandle legacy_ = nullptr;
  std::string dir_;
};

// ---------------------------------------------------------------------------- prefix index
class GpuGlobalKVCacheIndex {
 public:
  explicit GpuGlobalKVCacheIndex(xllm_ingest_t h) : h_(h) {}

  int instance_id(const std::string& name) {  // names -> bit positions (<= 64 instances)
    std::lock_guard<std::mutex> l(mu_);
    auto it = ids_.find(name);
    if (it != ids_.end()) return it->second;
    CHECK_LT(names_.size(), (size_t)XLLM_MAX_INSTANCES);
    ids_[name] = (int)names_.size();
    names_.push_back(name);
    return (int)names_.size() - 1;
  }
  const std::string& name_of(int id) const { return names_[id]; }

  // GlobalKVCacheMgr::record_updated_kvcaches (global_kvcache_mgr.cpp:177-225)
  void record_updated_kvcaches(const std::string& instance_name, const proto::KvCacheEvent& e) {
    auto pack = [](const google::protobuf::RepeatedPtrField<std::string>& f) {
      std::string out;
      for (const auto& k : f) out.append(k.data(), 16);
      return out;
    };
    const std::string s = pack(e.stored_cache()), o = pack(e.offload_cache()), r = pack(e.removed_cache());
    xllm_index_apply(h_, instance_id(instance_name), reinterpret_cast<const uint8_t*>(s.data()), s.size() / 16,
                     reinterpret_cast<const uint8_t*>(o.data()), o.size() / 16,
                     reinterpret_cast<const uint8_t*>(r.data()), r.size() / 16);
  }
  // GlobalKVCacheMgr::upload_kvcache's local effect (:227-247); the etcd write stays where it is
  bool upload_kvcache() { return xllm_index_publish(h_) == XLLM_OK; }

  // The table as the pairs the master keeps under XLLM:CACHE: (etcd_client.cpp:122-137), e.g. to seed a new etcd
  // cluster or to hand the index to a freshly elected master; host/index_snapshot.h.
  bool snapshot(const std::string& namespace_prefix, std::vector<xllm_host::CacheKv>* out) {
    std::lock_guard<std::mutex> l(mu_);
    return xllm_host::snapshot_index(h_, namespace_prefix, names_, out) == XLLM_OK;
  }
  // The constructor's start-up load (:47-51) and the replica watch (:133-175): the pairs of one listing / one
  // watch response (empty value = DELETE).  prefix_len = size of namespace + "XLLM:CACHE:".
  bool apply_etcd_pairs(const std::vector<xllm_host::CacheKv>& kvs, size_t prefix_len) {
    size_t skipped = 0;
    const int rc = xllm_host::apply_etcd_pairs(
        h_, prefix_len, kvs, [this](const std::string& n) { return instance_id(n); }, &skipped);
    if (skipped) LOG(ERROR) << skipped << " XLLM:CACHE pairs could not be parsed";
    return rc == XLLM_OK;
  }

  // GlobalKVCacheMgr::match (:73-131) for one request (the batch path goes through IngestBatcher)
  void match(const Slice<int32_t>& token_ids, int32_t block_size, OverlapScores* out) {
    const size_t nb = token_ids.size() / block_size;
    if (nb == 0) return;
    std::vector<uint8_t> keys(16 * nb);
    const int64_t zero = 0;
    const int32_t n_tok = (int32_t)token_ids.size(), n_blk = (int32_t)nb;
    xllm_hash_blocks(h_, 1, token_ids.data(), n_tok, &zero, &n_tok, keys.data(), (int64_t)nb, &zero);
    xllm_match_out m{};
    xllm_match_route(h_, 1, keys.data(), (int64_t)nb, &zero, &n_blk, &m, nullptr);
    out->max_block_num = m.max_block_num;
    out->max_matched_block_num = m.max_matched_block_num;
    for (int i = 0; i < XLLM_MAX_INSTANCES; ++i) {
      if (!((m.instances >> i) & 1)) continue;
      out->instances.insert(names_[i]);
      if (m.hbm_instance_score[i]) out->hbm_instance_score[names_[i]] = m.hbm_instance_score[i];
      if (m.dram_instance_score[i]) out->dram_instance_score[names_[i]] = m.dram_instance_score[i];
      if (m.ssd_instance_score[i]) out->ssd_instance_score[names_[i]] = m.ssd_instance_score[i];
    }
  }

 private:
  xllm_ingest_t h_;
  std::mutex mu_;
  std::unordered_map<std::string, int> ids_;
  std::vector<std::string> names_;
};

}  // namespace xllm_service
