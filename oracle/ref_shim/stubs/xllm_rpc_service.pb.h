// TEST INFRASTRUCTURE.  The accessors protoc generates for the messages the index write path reads
// (proto/xllm_rpc_service.proto:48-57: KvCacheEvent {repeated bytes stored_cache = 1, removed_cache = 2,
// offload_cache = 3}; LoadMetrics), hand-written because protoc is not in this image.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
namespace xllm_service { namespace proto {
class KvCacheEvent {
 public:
  int stored_cache_size() const { return (int)stored_.size(); }
  const std::string& stored_cache(int i) const { return stored_[i]; }
  void add_stored_cache(std::string v) { stored_.push_back(std::move(v)); }
  int removed_cache_size() const { return (int)removed_.size(); }
  const std::string& removed_cache(int i) const { return removed_[i]; }
  void add_removed_cache(std::string v) { removed_.push_back(std::move(v)); }
  int offload_cache_size() const { return (int)offload_.size(); }
  const std::string& offload_cache(int i) const { return offload_[i]; }
  void add_offload_cache(std::string v) { offload_.push_back(std::move(v)); }
 private:
  std::vector<std::string> stored_, removed_, offload_;
};
class LoadMetrics {
 public:
  uint64_t waiting_requests_num() const { return w_; }
  float gpu_cache_usage_perc() const { return u_; }
  uint64_t w_ = 0;
  float u_ = 0;
};
class LatencyMetrics {};
}}  // namespace xllm_service::proto
