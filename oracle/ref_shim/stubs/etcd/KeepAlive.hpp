// TEST INFRASTRUCTURE.  etcd::KeepAlive: leases never expire in the in-memory store.
#pragma once
#include "etcd/SyncClient.hpp"
namespace etcd {
class KeepAlive {
 public:
  KeepAlive(SyncClient&, int /*ttl*/) {}
  int64_t Lease() const { return 1; }
  void Cancel() {}
};
}  // namespace etcd
