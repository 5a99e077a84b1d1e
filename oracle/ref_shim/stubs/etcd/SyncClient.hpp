// TEST INFRASTRUCTURE (oracle/_ref build only).  An in-memory stand-in for etcd-cpp-apiv3 — just the calls the
// reference's scheduler/etcd_client/etcd_client.{h,cpp} makes — so that file and GlobalKVCacheMgr compile
// unmodified and run against a process-local key/value store: a master's upload_kvcache() really writes
// "XLLM:CACHE:"+key -> CacheLocations JSON pairs, and a replica's watch really receives PUT / DELETE events.
// One store per address string; watch callbacks are delivered synchronously by the writer (or once per
// begin_batch()/end_batch() bracket, like one watch response carrying several events).
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

// scheduler/etcd_client/etcd_client.h:63 names `unordered_map` unqualified; with the real etcd-cpp-apiv3 headers a
// using-declaration reaches it through cpprestsdk.  Reproduced here so the reference header compiles as is.
using std::unordered_map;

namespace etcd {

class Value {
 public:
  Value() = default;
  Value(std::string k, std::string v) : k_(std::move(k)), v_(std::move(v)) {}
  const std::string& key() const { return k_; }
  const std::string& as_string() const { return v_; }
 private:
  std::string k_, v_;
};

class Event {
 public:
  enum class EventType { PUT, DELETE_, INVALID };
  Event() = default;
  Event(EventType t, Value kv, bool has_prev, Value prev)
      : t_(t), kv_(std::move(kv)), has_prev_(has_prev), prev_(std::move(prev)) {}
  EventType event_type() const { return t_; }
  bool has_kv() const { return true; }
  const Value& kv() const { return kv_; }
  bool has_prev_kv() const { return has_prev_; }
  const Value& prev_kv() const { return prev_; }
 private:
  EventType t_ = EventType::INVALID;
  Value kv_;
  bool has_prev_ = false;
  Value prev_;
};

class Response {
 public:
  Response() = default;
  bool is_ok() const { return ok_; }
  std::string error_message() const { return ok_ ? "" : "fake etcd: not found"; }
  const Value& value() const { return value_; }
  const std::vector<std::string>& keys() const { return keys_; }
  const std::string& key(int i) const { return keys_[i]; }
  const Value& value(int i) const { return values_[i]; }
  const std::vector<Event>& events() const { return events_; }
  // filled by the fake
  bool ok_ = true;
  Value value_;
  std::vector<std::string> keys_;
  std::vector<Value> values_;
  std::vector<Event> events_;
};

namespace fake {
struct WatchEntry {
  std::string prefix;
  std::function<void(Response)> cb;
  bool live = true;
};
struct Store {
  std::recursive_mutex mu;
  std::map<std::string, std::string> kv;
  std::vector<std::shared_ptr<WatchEntry>> watchers;
  int batching = 0;
  std::vector<Event> pending;

  void emit(Event e) {
    pending.push_back(std::move(e));
    if (!batching) flush();
  }
  void flush() {
    std::vector<Event> evs;
    evs.swap(pending);
    if (evs.empty()) return;
    auto ws = watchers;   // a callback may add / cancel watchers
    for (auto& w : ws) {
      if (!w->live) continue;
      Response r;
      for (auto& e : evs)
        if (e.kv().key().compare(0, w->prefix.size(), w->prefix) == 0) r.events_.push_back(e);
      if (!r.events_.empty()) w->cb(r);
    }
  }
  void put(const std::string& k, const std::string& v) {
    std::lock_guard<std::recursive_mutex> g(mu);
    auto it = kv.find(k);
    bool had = it != kv.end();
    Value prev = had ? Value(k, it->second) : Value();
    kv[k] = v;
    emit(Event(Event::EventType::PUT, Value(k, v), had, prev));
  }
  bool rm(const std::string& k) {
    std::lock_guard<std::recursive_mutex> g(mu);
    auto it = kv.find(k);
    if (it == kv.end()) return false;
    Value prev(k, it->second);
    kv.erase(it);
    emit(Event(Event::EventType::DELETE_, Value(k, ""), true, prev));
    return true;
  }
  void begin_batch() {
    std::lock_guard<std::recursive_mutex> g(mu);
    ++batching;
  }
  void end_batch() {
    std::lock_guard<std::recursive_mutex> g(mu);
    if (--batching == 0) flush();
  }
};
inline std::shared_ptr<Store> store_for(const std::string& addr) {
  static std::mutex mu;
  static std::map<std::string, std::shared_ptr<Store>> stores;
  std::lock_guard<std::mutex> g(mu);
  auto& s = stores[addr];
  if (!s) s = std::make_shared<Store>();
  return s;
}
}  // namespace fake
}  // namespace etcd

namespace etcdv3 {
enum class CompareResult { EQUAL, GREATER, LESS, NOT_EQUAL };
class Transaction {
 public:
  void add_compare_create(const std::string& k, int64_t) { must_be_absent_.push_back(k); }
  void add_compare_version(const std::string&, CompareResult, int64_t) {}
  void add_success_put(const std::string& k, const std::string& v, int64_t = 0) { puts_.emplace_back(k, v); }
  void add_success_delete(const std::string& k) { dels_.push_back(k); }
  std::vector<std::string> must_be_absent_, dels_;
  std::vector<std::pair<std::string, std::string>> puts_;
};
}  // namespace etcdv3

namespace etcd {
class SyncClient {
 public:
  explicit SyncClient(const std::string& addr) : store_(fake::store_for(addr)) {}
  SyncClient(const std::string& addr, const std::string&, const std::string&) : store_(fake::store_for(addr)) {}
  Response put(const std::string& k, const std::string& v) {
    store_->put(k, v);
    return Response();
  }
  // Deleting an absent key: a healthy etcd answers OK with deleted = 0.  etcd-cpp-apiv3 turns that reply into
  // error 100 "key not found" (AsyncDeleteResponse), which makes upload_kvcache() return false whenever a block was
  // stored and removed inside one flush window and leaves its moved-from staging map behind
  // (global_kvcache_mgr.cpp:233-246).  That is an error path of a third-party client, outside the hot path's
  // contract; the fake answers OK unless rm_missing_is_error() is switched on (used by one test to show the effect).
  static bool& rm_missing_is_error() {
    static bool v = false;
    return v;
  }
  Response rm(const std::string& k) {
    Response r;
    r.ok_ = store_->rm(k) || !rm_missing_is_error();
    return r;
  }
  Response get(const std::string& k) {
    std::lock_guard<std::recursive_mutex> g(store_->mu);
    Response r;
    auto it = store_->kv.find(k);
    r.ok_ = it != store_->kv.end();
    if (r.ok_) r.value_ = Value(k, it->second);
    return r;
  }
  Response ls(const std::string& prefix) {
    std::lock_guard<std::recursive_mutex> g(store_->mu);
    Response r;
    for (auto it = store_->kv.lower_bound(prefix); it != store_->kv.end(); ++it) {
      if (it->first.compare(0, prefix.size(), prefix) != 0) break;
      r.keys_.push_back(it->first);
      r.values_.emplace_back(it->first, it->second);
    }
    return r;
  }
  Response txn(const etcdv3::Transaction& t) {
    std::lock_guard<std::recursive_mutex> g(store_->mu);
    Response r;
    for (auto& k : t.must_be_absent_)
      if (store_->kv.count(k)) r.ok_ = false;
    if (!r.ok_) return r;
    store_->begin_batch();
    for (auto& p : t.puts_) store_->put(p.first, p.second);
    for (auto& k : t.dels_) store_->rm(k);
    store_->end_batch();
    return r;
  }
  std::shared_ptr<fake::Store> store_;
};
}  // namespace etcd
