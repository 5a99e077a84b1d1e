// TEST INFRASTRUCTURE.  The slice of absl::Mutex that common/concurrent_queue.h uses (MutexLock, Await with a
// Condition over a lambda), on std::mutex + condition_variable: every unlock notifies, Await re-checks.
#pragma once
#include <condition_variable>
#include <functional>
#include <mutex>
namespace absl {
class Condition {
 public:
  template <typename F>
  explicit Condition(const F* f) : fn_([f] { return (*f)(); }) {}
  bool eval() const { return fn_(); }
 private:
  std::function<bool()> fn_;
};
class Mutex {
 public:
  void Lock() { mu_.lock(); }
  void Unlock() {
    mu_.unlock();
    cv_.notify_all();
  }
  void Await(const Condition& c) {   // called with the mutex held; returns with it held
    std::unique_lock<std::mutex> l(mu_, std::adopt_lock);
    while (!c.eval()) cv_.wait(l);
    l.release();
  }
 private:
  std::mutex mu_;
  std::condition_variable cv_;
};
class MutexLock {
 public:
  explicit MutexLock(Mutex* m) : m_(m) { m_->Lock(); }
  ~MutexLock() { m_->Unlock(); }
 private:
  Mutex* m_;
};
}  // namespace absl
