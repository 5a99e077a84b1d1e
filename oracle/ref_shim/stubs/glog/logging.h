// TEST INFRASTRUCTURE (oracle/_ref build only).  A minimal stand-in for <glog/logging.h>, enough for the
// reference's own translation units to compile unmodified: LOG / DLOG / VLOG streams (dropped unless
// XLLM_REF_LOG is set; FATAL aborts) and the CHECK family (abort with a message when violated).
#pragma once
// The real glog headers also drag in <atomic>, <cstring>, ... which some reference files rely on transitively.
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <iostream>
#include <sstream>
#include <string>

namespace ref_glog {
struct Sink {
  bool fatal;
  std::ostringstream os;
  explicit Sink(bool f) : fatal(f) {}
  ~Sink() {
    if (fatal) {
      std::cerr << "[ref FATAL] " << os.str() << std::endl;
      std::abort();
    }
    if (std::getenv("XLLM_REF_LOG")) std::cerr << "[ref] " << os.str() << std::endl;
  }
  template <typename T>
  Sink& operator<<(const T& v) {
    os << v;
    return *this;
  }
  Sink& operator<<(std::ostream& (*m)(std::ostream&)) {
    os << m;
    return *this;
  }
};
struct Voidify {
  void operator&(const Sink&) {}
};
}  // namespace ref_glog

#define REF_GLOG_SEV_INFO false
#define REF_GLOG_SEV_WARNING false
#define REF_GLOG_SEV_ERROR false
#define REF_GLOG_SEV_FATAL true
#define LOG(sev) ::ref_glog::Sink(REF_GLOG_SEV_##sev)
#define DLOG(sev) LOG(sev)
#define VLOG(n) ::ref_glog::Sink(false)
#define LOG_IF(sev, cond) !(cond) ? (void)0 : ::ref_glog::Voidify() & LOG(sev)
#define LOG_EVERY_N(sev, n) LOG(sev)
#define LOG_FIRST_N(sev, n) LOG(sev)
#define CHECK(cond) (cond) ? (void)0 : ::ref_glog::Voidify() & ::ref_glog::Sink(true) << "Check failed: " #cond " "
#define REF_GLOG_CHECK_OP(a, b, op) CHECK((a)op(b))
#define CHECK_EQ(a, b) REF_GLOG_CHECK_OP(a, b, ==)
#define CHECK_NE(a, b) REF_GLOG_CHECK_OP(a, b, !=)
#define CHECK_LT(a, b) REF_GLOG_CHECK_OP(a, b, <)
#define CHECK_LE(a, b) REF_GLOG_CHECK_OP(a, b, <=)
#define CHECK_GT(a, b) REF_GLOG_CHECK_OP(a, b, >)
#define CHECK_GE(a, b) REF_GLOG_CHECK_OP(a, b, >=)
#define DCHECK(cond) CHECK(cond)
#define CHECK_NOTNULL(p) (p)
