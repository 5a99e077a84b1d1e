// TEST INFRASTRUCTURE.  etcd::Watcher over the in-memory store of SyncClient.hpp.
#pragma once
#include "etcd/SyncClient.hpp"
namespace etcd {
class Watcher {
 public:
  Watcher(SyncClient& c, const std::string& prefix, std::function<void(Response)> cb, bool /*recursive*/ = true)
      : store_(c.store_), entry_(std::make_shared<fake::WatchEntry>()) {
    entry_->prefix = prefix;
    entry_->cb = std::move(cb);
    std::lock_guard<std::recursive_mutex> g(store_->mu);
    store_->watchers.push_back(entry_);
  }
  ~Watcher() { Cancel(); }
  bool Cancel() {
    std::lock_guard<std::recursive_mutex> g(store_->mu);
    entry_->live = false;
    for (size_t i = 0; i < store_->watchers.size(); ++i)
      if (store_->watchers[i] == entry_) {
        store_->watchers.erase(store_->watchers.begin() + i);
        break;
      }
    return true;
  }
 private:
  std::shared_ptr<fake::Store> store_;
  std::shared_ptr<fake::WatchEntry> entry_;
};
}  // namespace etcd
