// TEST INFRASTRUCTURE.  request/request.h pulls in brpc, minja and protobuf; CacheAwareRouting
// (scheduler/loadbalance_policy/cache_aware_routing.cpp) touches two members of Request only
// (request/request.h:61 token_ids, :64 routing), so the _ref build substitutes this two-member struct.
#pragma once
#include <cstdint>
#include <memory>
#include <vector>
#include "common/types.h"
namespace xllm_service {
struct Request {
  bool offline = false;            // request/request.h:41
  std::vector<int32_t> token_ids;  // request/request.h:61
  Routing routing;                 // request/request.h:64
};
}  // namespace xllm_service
