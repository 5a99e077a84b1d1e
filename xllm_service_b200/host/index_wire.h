// index_wire.h — the prefix index's wire / disk form, header-only C++17, no dependencies.
//
// What the reference keeps in etcd for every block key:
//   key    namespace + "XLLM:CACHE:" + the 16 raw key bytes
//          (global_kvcache_mgr.cpp:27 ETCD_CACHE_PREFIX, etcd_client.cpp:122-137 EtcdClient::set,
//           hash_util.h:27-30 XXH3Key::to_string, utils.cpp:125-132 build_etcd_key_with_namespace)
//   value  CacheLocations::serialize_to_json().dump()                       (common/types.h:325-331)
//          = nlohmann's compact dump of an object whose keys nlohmann keeps sorted:
//          {"dram_instance_set":[...],"hbm_instance_set":[...],"ssd_instance_set":[...]}
//          with the instance names as JSON strings; an EMPTY CacheLocations is a delete (etcd_client.cpp:128-129)
//   read   CacheLocations::parse_from_json (types.h:335-357): json.at(k).get<vector<string>>() for the three keys —
//          any key order / whitespace / extra keys accepted, a missing key or a non-string element rejects the entry
// Here the three sets are 64-bit instance masks (bit = instance id of the adaptor's name table), so this header
// converts masks <-> that JSON and block keys <-> etcd keys.  The name order inside an array is the one thing the
// reference does not define (unordered_set iteration); this writer emits ascending instance ids.
// Pinned against the real nlohmann::json by tests/test_index_wire.py (tests/cpp/index_wire_main.cc).
#pragma once
#include <stdint.h>
#include <string.h>

#include <functional>
#include <string>
#include <vector>

namespace xllm_host {

inline const char* etcd_cache_prefix() { return "XLLM:CACHE:"; }

// namespace prefix (may be empty) + "XLLM:CACHE:" + 16 raw bytes
inline std::string cache_etcd_key(const std::string& namespace_prefix, const uint8_t key16[16]) {
  std::string k = namespace_prefix;
  k += etcd_cache_prefix();
  k.append(reinterpret_cast<const char*>(key16), 16);
  return k;
}
// The reference cuts the watched prefix off and memcpy's 16 bytes (etcd_client.cpp:185, hash_util.h:23-25).
inline bool parse_cache_etcd_key(const std::string& full_key, size_t prefix_len, uint8_t key16[16]) {
  if (full_key.size() < prefix_len + 16) return false;
  memcpy(key16, full_key.data() + prefix_len, 16);
  return true;
}

namespace wire_detail {

// nlohmann::detail::serializer::dump_escaped with ensure_ascii = false, error_handler strict
inline bool append_json_string(const std::string& s, std::string* out) {
  static const char* hex = "0123456789abcdef";
  out->push_back('"');
  const unsigned char* p = reinterpret_cast<const unsigned char*>(s.data());
  const size_t n = s.size();
  for (size_t i = 0; i < n;) {
    const unsigned char c = p[i];
    if (c < 0x80) {
      switch (c) {
        case '"': *out += "\\\""; break;
        case '\\': *out += "\\\\"; break;
        case '\b': *out += "\\b"; break;
        case '\f': *out += "\\f"; break;
        case '\n': *out += "\\n"; break;
        case '\r': *out += "\\r"; break;
        case '\t': *out += "\\t"; break;
        default:
          if (c < 0x20) {
            *out += "\\u00";
            out->push_back(hex[c >> 4]);
            out->push_back(hex[c & 15]);
          } else {
            out->push_back((char)c);
          }
      }
      ++i;
      continue;
    }
    // multi-byte: must be well-formed UTF-8 (nlohmann throws type_error 316 otherwise)
    int len = 0;
    uint32_t cp = 0;
    if ((c & 0xE0) == 0xC0) { len = 2; cp = c & 0x1F; }
    else if ((c & 0xF0) == 0xE0) { len = 3; cp = c & 0x0F; }
    else if ((c & 0xF8) == 0xF0) { len = 4; cp = c & 0x07; }
    else return false;
    if (i + len > n) return false;
    for (int k = 1; k < len; ++k) {
      if ((p[i + k] & 0xC0) != 0x80) return false;
      cp = (cp << 6) | (p[i + k] & 0x3F);
    }
    if ((len == 2 && cp < 0x80) || (len == 3 && (cp < 0x800 || (cp >= 0xD800 && cp < 0xE000))) ||
        (len == 4 && (cp < 0x10000 || cp > 0x10FFFF)))
      return false;
    out->append(s, i, len);
    i += len;
  }
  out->push_back('"');
  return true;
}

inline bool append_name_array(uint64_t mask, const std::vector<std::string>& names, std::string* out) {
  out->push_back('[');
  bool first = true;
  for (int i = 0; i < 64; ++i) {
    if (!((mask >> i) & 1)) continue;
    if ((size_t)i >= names.size()) return false;
    if (!first) out->push_back(',');
    first = false;
    if (!append_json_string(names[(size_t)i], out)) return false;
  }
  out->push_back(']');
  return true;
}

// A small strict JSON reader: enough to walk any value, collecting the three arrays we need.
class Reader {
 public:
  Reader(const char* p, const char* e) : p_(p), e_(e) {}
  void ws() { while (p_ < e_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r')) ++p_; }
  bool at_end() { ws(); return p_ == e_; }
  bool lit(const char* s) {
    const size_t n = strlen(s);
    if ((size_t)(e_ - p_) < n || memcmp(p_, s, n) != 0) return false;
    p_ += n;
    return true;
  }
  static void put_utf8(uint32_t cp, std::string* s) {
    if (cp < 0x80) s->push_back((char)cp);
    else if (cp < 0x800) { s->push_back((char)(0xC0 | (cp >> 6))); s->push_back((char)(0x80 | (cp & 0x3F))); }
    else if (cp < 0x10000) { s->push_back((char)(0xE0 | (cp >> 12))); s->push_back((char)(0x80 | ((cp >> 6) & 0x3F))); s->push_back((char)(0x80 | (cp & 0x3F))); }
    else { s->push_back((char)(0xF0 | (cp >> 18))); s->push_back((char)(0x80 | ((cp >> 12) & 0x3F))); s->push_back((char)(0x80 | ((cp >> 6) & 0x3F))); s->push_back((char)(0x80 | (cp & 0x3F))); }
  }
  bool hex4(uint32_t* v) {
    if (e_ - p_ < 4) return false;
    *v = 0;
    for (int i = 0; i < 4; ++i) {
      const char c = *p_++;
      *v <<= 4;
      if (c >= '0' && c <= '9') *v |= (uint32_t)(c - '0');
      else if (c >= 'a' && c <= 'f') *v |= (uint32_t)(c - 'a' + 10);
      else if (c >= 'A' && c <= 'F') *v |= (uint32_t)(c - 'A' + 10);
      else return false;
    }
    return true;
  }
  bool string(std::string* s) {
    if (p_ >= e_ || *p_ != '"') return false;
    ++p_;
    while (p_ < e_ && *p_ != '"') {
      const unsigned char c = (unsigned char)*p_;
      if (c < 0x20) return false;  // control characters must be escaped
      if (c != '\\') { s->push_back(*p_++); continue; }
      if (++p_ >= e_) return false;
      switch (*p_++) {
        case '"': s->push_back('"'); break;
        case '\\': s->push_back('\\'); break;
        case '/': s->push_back('/'); break;
        case 'b': s->push_back('\b'); break;
        case 'f': s->push_back('\f'); break;
        case 'n': s->push_back('\n'); break;
        case 'r': s->push_back('\r'); break;
        case 't': s->push_back('\t'); break;
        case 'u': {
          uint32_t cp;
          if (!hex4(&cp)) return false;
          if (cp >= 0xD800 && cp < 0xDC00) {  // high surrogate: a low one must follow
            uint32_t lo;
            if (e_ - p_ < 6 || p_[0] != '\\' || p_[1] != 'u') return false;
            p_ += 2;
            if (!hex4(&lo) || lo < 0xDC00 || lo > 0xDFFF) return false;
            cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
          } else if (cp >= 0xDC00 && cp < 0xE000) {
            return false;
          }
          put_utf8(cp, s);
          break;
        }
        default: return false;
      }
    }
    if (p_ >= e_) return false;
    ++p_;
    return true;
  }
  // skips any JSON value
  bool skip(int depth = 0) {
    if (depth > 64) return false;
    ws();
    if (p_ >= e_) return false;
    if (*p_ == '"') { std::string t; return string(&t); }
    if (*p_ == '{' || *p_ == '[') {
      const char close = *p_ == '{' ? '}' : ']';
      const bool obj = *p_ == '{';
      ++p_;
      ws();
      if (p_ < e_ && *p_ == close) { ++p_; return true; }
      for (;;) {
        if (obj) {
          ws();
          std::string k;
          if (!string(&k)) return false;
          ws();
          if (p_ >= e_ || *p_++ != ':') return false;
        }
        if (!skip(depth + 1)) return false;
        ws();
        if (p_ < e_ && *p_ == ',') { ++p_; continue; }
        if (p_ < e_ && *p_ == close) { ++p_; return true; }
        return false;
      }
    }
    if (lit("true") || lit("false") || lit("null")) return true;
    const char* s = p_;
    if (p_ < e_ && *p_ == '-') ++p_;
    while (p_ < e_ && ((*p_ >= '0' && *p_ <= '9') || *p_ == '.' || *p_ == 'e' || *p_ == 'E' || *p_ == '+' || *p_ == '-')) ++p_;
    return p_ > s;
  }
  // ["a","b",...] -> strings; anything else fails (get<vector<string>> throws on non-array / non-string)
  bool string_array(std::vector<std::string>* out) {
    ws();
    if (p_ >= e_ || *p_ != '[') return false;
    ++p_;
    ws();
    if (p_ < e_ && *p_ == ']') { ++p_; return true; }
    for (;;) {
      ws();
      out->emplace_back();
      if (!string(&out->back())) return false;
      ws();
      if (p_ < e_ && *p_ == ',') { ++p_; continue; }
      if (p_ < e_ && *p_ == ']') { ++p_; return true; }
      return false;
    }
  }
  const char* p_;
  const char* e_;
};

}  // namespace wire_detail

// CacheLocations::serialize_to_json().dump() for the three masks.  names[i] = name of instance id i.
// false: a set bit without a name, or a name that is not valid UTF-8 (nlohmann would throw).
inline bool cache_locations_to_json(uint64_t hbm, uint64_t dram, uint64_t ssd, const std::vector<std::string>& names,
                                    std::string* out) {
  out->clear();
  *out += "{\"dram_instance_set\":";
  if (!wire_detail::append_name_array(dram, names, out)) return false;
  *out += ",\"hbm_instance_set\":";
  if (!wire_detail::append_name_array(hbm, names, out)) return false;
  *out += ",\"ssd_instance_set\":";
  if (!wire_detail::append_name_array(ssd, names, out)) return false;
  out->push_back('}');
  return true;
}

// CacheLocations::parse_from_json.  id_of(name) returns the instance id (0..63) — it may register a new name —
// or a negative number to reject the entry.  Duplicate keys: the last one wins, as in nlohmann.
inline bool cache_locations_from_json(const std::string& json, const std::function<int(const std::string&)>& id_of,
                                      uint64_t* hbm, uint64_t* dram, uint64_t* ssd) {
  wire_detail::Reader r(json.data(), json.data() + json.size());
  r.ws();
  if (r.p_ >= r.e_ || *r.p_ != '{') return false;
  ++r.p_;
  std::vector<std::string> sets[3];
  bool have[3] = {false, false, false};
  r.ws();
  if (r.p_ < r.e_ && *r.p_ == '}') {
    ++r.p_;
  } else {
    for (;;) {
      r.ws();
      std::string k;
      if (!r.string(&k)) return false;
      r.ws();
      if (r.p_ >= r.e_ || *r.p_++ != ':') return false;
      const int which = k == "hbm_instance_set" ? 0 : (k == "dram_instance_set" ? 1 : (k == "ssd_instance_set" ? 2 : -1));
      if (which < 0) {
        if (!r.skip()) return false;
      } else {
        sets[which].clear();
        // a value of another type still has to be well-formed JSON before it is rejected: both are failures here
        if (!r.string_array(&sets[which])) return false;
        have[which] = true;
      }
      r.ws();
      if (r.p_ < r.e_ && *r.p_ == ',') { ++r.p_; continue; }
      if (r.p_ < r.e_ && *r.p_ == '}') { ++r.p_; break; }
      return false;
    }
  }
  if (!r.at_end()) return false;
  if (!have[0] || !have[1] || !have[2]) return false;  // json.at(key) throws out_of_range
  uint64_t m[3] = {0, 0, 0};
  for (int w = 0; w < 3; ++w)
    for (const auto& name : sets[w]) {
      const int id = id_of(name);
      if (id < 0 || id >= 64) return false;
      m[w] |= 1ull << id;
    }
  *hbm = m[0];
  *dram = m[1];
  *ssd = m[2];
  return true;
}

}  // namespace xllm_host
