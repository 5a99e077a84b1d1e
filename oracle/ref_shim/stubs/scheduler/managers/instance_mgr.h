// TEST INFRASTRUCTURE.  scheduler/managers/instance_mgr.h needs brpc and Eigen.  The hot path uses one method,
// InstanceMgr::get_load_metrics (instance_mgr.cpp:287-359), over four members (instance_mgr.h:155-189).  This
// declaration keeps those members with the reference's names and types; the method BODY is not restated: the build
// script cuts instance_mgr.cpp:63-66 (is_instance_schedulable) and :287-359 out of the reference into a generated
// file under oracle/_ref/ at build time.
#pragma once
#include <shared_mutex>
#include <string>
#include <unordered_map>
#include "common/types.h"
#include "request/request.h"
namespace xllm_service {
class InstanceMgr final {
 public:
  InstanceMgr() = default;
  void get_load_metrics(LoadBalanceInfos* infos);
  // instance_mgr.h:155-189
  mutable std::shared_mutex cluster_mutex_;
  mutable std::shared_mutex metrics_mutex_;
  std::unordered_map<std::string, InstanceMetaInfo> instances_;
  std::unordered_map<std::string, LoadMetrics> load_metrics_;
};
}  // namespace xllm_service
