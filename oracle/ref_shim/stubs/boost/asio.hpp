// TEST INFRASTRUCTURE.  common/utils.cpp uses boost::asio only in get_local_ip(), which the hot path never calls;
// these empty shells let the file compile unmodified (normalize_etcd_namespace / build_etcd_key_with_namespace
// are what the index's wire form needs from it).
#pragma once
#include <string>
namespace boost { namespace asio {
struct io_service {};
namespace ip {
inline std::string host_name() { return "localhost"; }
struct address {
  bool is_loopback() const { return true; }
  bool is_v4() const { return true; }
  std::string to_string() const { return "127.0.0.1"; }
};
struct tcp {
  struct endpoint_t { ip::address address() const { return {}; } };
  struct entry { endpoint_t endpoint() const { return {}; } };
  struct resolver {
    explicit resolver(io_service&) {}
    struct query { query(const std::string&, const std::string&) {} };
    struct iterator {
      bool end = true;
      const entry* operator->() const { static entry e; return &e; }
      iterator& operator++() { end = true; return *this; }
      bool operator!=(const iterator& o) const { return end != o.end; }
    };
    iterator resolve(const query&) { return {}; }
  };
};
}  // namespace ip
}}  // namespace boost::asio
