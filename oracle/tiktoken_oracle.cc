// oracle/tiktoken_oracle.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the reference's in-repo tiktoken backend as the service actually runs it:
//   xllm_service/tokenizer/tiktoken_tokenizer.cpp:115-153  load_vocab: lines "base64(token) SP rank"
//   xllm_service/tokenizer/tiktoken_tokenizer.cpp:155-234  byte_pair_encode: parts = (start, rank of the pair
//        starting there); repeatedly merge the MIN rank, leftmost (:197-207); re-rank the two neighbours
//        (:215-218); emit encoder_[bytes] per part, parts without an entry are logged and skipped (:222-233)
//   xllm_service/tokenizer/tiktoken_tokenizer.cpp:236-254  encode_internal: the service never sets `pattern`
//        (tokenizer_args.cpp:30-71 does not load it) => regex_ == nullptr => the WHOLE text is one piece
//   xllm_service/tokenizer/tiktoken_tokenizer.cpp:256-294  encode: no prefix tokens / special-token regex in the
//        service => encode == encode_internal(text), appending
// Pinned against pip tiktoken 0.12.0 (tiktoken.Encoding._encode_single_piece on the same ranks) by
// tests/test_oracle_tiktoken.py and the committed vectors tests/golden/tiktoken_goldens.json.
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <limits>
#include <optional>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

namespace {

struct Tik {
  std::unordered_map<std::string, int32_t> encoder_;
};

bool b64_decode(const std::string& in, std::string* out) {
  static int8_t T[256];
  static bool init = false;
  if (!init) {
    memset(T, -1, sizeof(T));
    const char* a = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    for (int i = 0; i < 64; ++i) T[(uint8_t)a[i]] = (int8_t)i;
    init = true;
  }
  out->clear();
  uint32_t acc = 0;
  int bits = 0;
  for (char c : in) {
    if (c == '=') break;
    const int8_t v = T[(uint8_t)c];
    if (v < 0) return false;
    acc = (acc << 6) | (uint32_t)v;
    bits += 6;
    if (bits >= 8) {
      bits -= 8;
      out->push_back((char)((acc >> bits) & 0xFF));
    }
  }
  return true;
}

// tiktoken_tokenizer.cpp:155-234
void byte_pair_encode(const Tik& t, std::string_view piece, std::vector<int32_t>* ids) {
  if (piece.empty()) return;
  std::vector<std::pair<int32_t, int32_t>> parts;
  parts.reserve(piece.size() + 1);
  const int32_t kMaxRank = std::numeric_limits<int32_t>::max();
  for (int32_t i = 0; i <= (int32_t)piece.size(); ++i) parts.emplace_back(i, kMaxRank);
  auto get_rank = [&](int32_t start, int32_t skip) -> std::optional<int32_t> {
    if (start + skip + 2 < (int32_t)parts.size()) {
      const auto s = parts[start].first;
      const auto e = parts[start + skip + 2].first;
      auto it = t.encoder_.find(std::string(piece.substr(s, e - s)));
      if (it != t.encoder_.end()) return it->second;
    }
    return std::nullopt;
  };
  for (int32_t i = 0; i < (int32_t)parts.size() - 2; ++i) {
    const auto rank = get_rank(i, 0);
    if (rank.has_value()) parts[i].second = rank.value();
  }
  while (parts.size() > 1) {
    int32_t min_rank = kMaxRank, min_i = 0;
    for (int32_t i = 0; i < (int32_t)parts.size() - 1; ++i) {
      if (parts[i].second < min_rank) {
        min_rank = parts[i].second;
        min_i = i;
      }
    }
    if (min_rank == kMaxRank) break;
    parts[min_i].second = get_rank(min_i, 1).value_or(kMaxRank);
    if (min_i > 0) parts[min_i - 1].second = get_rank(min_i - 1, 1).value_or(kMaxRank);
    parts.erase(parts.begin() + min_i + 1);
  }
  for (int32_t i = 0; i < (int32_t)parts.size() - 1; ++i) {
    const auto s = parts[i].first, e = parts[i + 1].first;
    auto it = t.encoder_.find(std::string(piece.substr(s, e - s)));
    if (it != t.encoder_.end()) ids->push_back(it->second);  // else: LOG(ERROR) and skip (:228-229)
  }
}

}  // namespace

extern "C" {

void* oracle_tik_load(const char* vocab_file, char* err, size_t cap) {
  FILE* f = fopen(vocab_file, "rb");
  if (!f) {
    if (err) snprintf(err, cap, "Failed to open vocab file: %s", vocab_file);
    return nullptr;
  }
  Tik* t = new Tik();
  char line[1 << 16];
  while (fgets(line, sizeof(line), f)) {
    std::string l(line);
    while (!l.empty() && (l.back() == '\n' || l.back() == '\r')) l.pop_back();
    if (l.empty()) continue;
    const size_t sp = l.find(' ');
    if (sp == std::string::npos || l.find(' ', sp + 1) != std::string::npos) continue;  // parts.size() != 2
    std::string tok;
    if (!b64_decode(l.substr(0, sp), &tok)) continue;
    char* endp = nullptr;
    const long rank = strtol(l.c_str() + sp + 1, &endp, 10);
    if (endp == l.c_str() + sp + 1 || *endp != '\0') continue;
    t->encoder_.try_emplace(tok, (int32_t)rank);
  }
  fclose(f);
  return t;
}
void oracle_tik_free(void* h) { delete (Tik*)h; }
long oracle_tik_vocab_size(void* h) { return (long)((Tik*)h)->encoder_.size(); }
// TiktokenTokenizer::encode as configured by the service; returns the id count (only `cap` are written).
long oracle_tik_encode(void* h, const char* text, size_t len, int32_t* out, size_t cap) {
  std::vector<int32_t> ids;
  byte_pair_encode(*(Tik*)h, std::string_view(text, len), &ids);
  memcpy(out, ids.data(), sizeof(int32_t) * (ids.size() < cap ? ids.size() : cap));
  return (long)ids.size();
}

}  // extern "C"
