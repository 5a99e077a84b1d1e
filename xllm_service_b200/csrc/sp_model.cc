// sp_model.cc — SentencePiece `.model` -> device tables (host side, no CUDA calls).
//
// The reference delegates all of this to libsentencepiece behind
// SentencePieceTokenizer (xllm_service/tokenizer/sentencepiece_tokenizer.cpp:43-54,115-128).
// The tables built here encode the same decisions libsentencepiece makes at encode time
// for a BPE model:
//   * pieces_ / reserved_id_map_ split and PieceToId order (model_interface.cc)
//   * BPE merge legality: "A+B may merge iff the string A||B is a NORMAL piece", priority
//     = piece score (higher first), leftmost on ties (bpe_model.cc)
//   * byte fallback / unknown handling (sentencepiece_processor.cc)
//   * the normalizer's precompiled charsmap (normalizer.cc)
#include "sp_model.h"

#include <stdio.h>
#include <string.h>
#include <sys/stat.h>

#include <algorithm>
#include <map>
#include <string_view>
#include <unordered_map>

#include "common.cuh"

namespace xllm {
namespace {

enum PieceType { kNormal = 1, kUnknown = 2, kControl = 3, kUserDefined = 4, kUnused = 5, kByte = 6 };

// Minimal protobuf wire-format cursor (varint / fixed32 / fixed64 / length-delimited).
class Wire {
 public:
  Wire(const char* d, size_t n) : p_((const uint8_t*)d), e_((const uint8_t*)d + n) {}
  bool good() const { return good_; }
  bool more() const { return good_ && p_ < e_; }
  bool field(uint32_t* num, uint32_t* type, uint64_t* scalar, std::string_view* span) {
    uint64_t tag;
    if (!more() || !varint(&tag)) return false;
    *num = (uint32_t)(tag >> 3);
    *type = (uint32_t)(tag & 7);
    if (*type == 0) return varint(scalar);
    if (*type == 1) return fixed(8, scalar);
    if (*type == 5) return fixed(4, scalar);
    if (*type == 2) {
      uint64_t n;
      if (!varint(&n) || n > (uint64_t)(e_ - p_)) return fail();
      *span = std::string_view((const char*)p_, (size_t)n);
      p_ += n;
      return true;
    }
    return fail();
  }

 private:
  bool fail() { good_ = false; return false; }
  bool varint(uint64_t* v) {
    *v = 0;
    for (int s = 0; s < 64 && p_ < e_; s += 7) {
      const uint8_t b = *p_++;
      *v |= (uint64_t)(b & 0x7F) << s;
      if (b < 0x80) return true;
    }
    return fail();
  }
  bool fixed(int n, uint64_t* v) {
    if (e_ - p_ < n) return fail();
    *v = 0;
    memcpy(v, p_, n);
    p_ += n;
    return true;
  }
  const uint8_t* p_;
  const uint8_t* e_;
  bool good_ = true;
};

struct RawPiece {
  std::string s;
  float score = 0.f;
  int type = kNormal;
};
struct RawModel {
  std::vector<RawPiece> pieces;
  int model_type = 1;  // UNIGRAM is the proto default
  bool byte_fallback = false, suffix_ws = false;
  std::string charsmap;
  bool dummy_prefix = true, remove_ws = true, escape_ws = true;
};

bool parse(const std::string& blob, RawModel* m) {
  Wire top(blob.data(), blob.size());
  uint32_t num, type;
  uint64_t val;
  std::string_view span;
  while (top.field(&num, &type, &val, &span)) {
    if (type != 2) continue;
    Wire sub(span.data(), span.size());
    uint32_t n2, t2;
    uint64_t v2;
    std::string_view s2;
    if (num == 1) {
      RawPiece pc;
      while (sub.field(&n2, &t2, &v2, &s2)) {
        if (n2 == 1 && t2 == 2) pc.s.assign(s2);
        if (n2 == 2 && t2 == 5) { uint32_t bits = (uint32_t)v2; memcpy(&pc.score, &bits, 4); }
        if (n2 == 3 && t2 == 0) pc.type = (int)v2;
      }
      m->pieces.push_back(std::move(pc));
    } else if (num == 2) {
      while (sub.field(&n2, &t2, &v2, &s2)) {
        if (t2 != 0) continue;
        if (n2 == 3) m->model_type = (int)v2;
        if (n2 == 35) m->byte_fallback = v2 != 0;
        if (n2 == 24) m->suffix_ws = v2 != 0;
      }
    } else if (num == 3) {
      while (sub.field(&n2, &t2, &v2, &s2)) {
        if (n2 == 2 && t2 == 2) m->charsmap.assign(s2);
        if (n2 == 3 && t2 == 0) m->dummy_prefix = v2 != 0;
        if (n2 == 4 && t2 == 0) m->remove_ws = v2 != 0;
        if (n2 == 5 && t2 == 0) m->escape_ws = v2 != 0;
      }
    }
    if (!sub.good()) return false;
  }
  return top.good();
}

// Splits a UTF-8 string at char boundaries (lead-byte lengths, as bpe_model.cc's OneCharLen split).
std::vector<size_t> char_bounds(const std::string& s) {
  static const char kLen[] = "\1\1\1\1\1\1\1\1\1\1\1\1\2\2\3\4";
  std::vector<size_t> b{0};
  size_t i = 0;
  while (i < s.size()) {
    i = std::min(s.size(), i + (size_t)kLen[((uint8_t)s[i]) >> 4]);
    b.push_back(i);
  }
  return b;
}

// Decodes one well-formed UTF-8 char; returns false for anything that is not exactly one canonical char.
bool single_cp(const std::string& s, uint32_t* cp) {
  const uint8_t* b = (const uint8_t*)s.data();
  const size_t n = s.size();
  if (n == 1 && b[0] < 0x80) { *cp = b[0]; return true; }
  if (n == 2 && (b[0] & 0xE0) == 0xC0 && (b[1] & 0xC0) == 0x80) { *cp = ((b[0] & 0x1F) << 6) | (b[1] & 0x3F); return *cp >= 0x80; }
  if (n == 3 && (b[0] & 0xF0) == 0xE0 && (b[1] & 0xC0) == 0x80 && (b[2] & 0xC0) == 0x80) {
    *cp = ((b[0] & 0x0F) << 12) | ((b[1] & 0x3F) << 6) | (b[2] & 0x3F);
    return *cp >= 0x800 && !(*cp >= 0xD800 && *cp < 0xE000);
  }
  if (n == 4 && (b[0] & 0xF8) == 0xF0 && (b[1] & 0xC0) == 0x80 && (b[2] & 0xC0) == 0x80 && (b[3] & 0xC0) == 0x80) {
    *cp = ((b[0] & 0x07) << 18) | ((b[1] & 0x3F) << 12) | ((b[2] & 0x3F) << 6) | (b[3] & 0x3F);
    return *cp >= 0x10000 && *cp <= 0x10FFFF;
  }
  return false;
}

inline uint32_t pow2_at_least(size_t n) {
  uint32_t p = 16;
  while (p < n) p <<= 1;
  return p;
}

}  // namespace

// Same mixers as the device side (sp_encode.cu); kept in one place via this header-less contract:
// Multiplicative hash of a symbol pair; the table index is its TOP log2(slots) bits (3 instructions on
// the device: IMAD, IMUL, SHF).  The caller masks with slots - 1 after shifting.
uint32_t sp_hash_pair(uint32_t a, uint32_t b) { return (a * 0x9E3779B1u + b) * 0x85EBCA6Bu; }
uint32_t sp_pair_slot(uint32_t a, uint32_t b, uint32_t n_slots) {
  uint32_t lg = 0;
  while ((1u << lg) < n_slots) ++lg;
  return lg ? (sp_hash_pair(a, b) >> (32 - lg)) : 0;
}
uint32_t sp_hash_cp(uint32_t cp) {
  uint32_t h = cp * 0x9E3779B1u;
  h ^= h >> 16;
  return h;
}

uint32_t sp_pair_slot_fwd(uint32_t a, uint32_t b, uint32_t n) { return sp_pair_slot(a, b, n); }

// slot of (parent node, byte) in SpTables::uni_trie; the device uses the same arithmetic (sp_encode.cu uni_trie_step)
uint32_t sp_trie_slot(uint32_t parent, uint32_t byte, uint32_t n_slots) {
  uint32_t h = (parent * 256u + byte) * 0x9E3779B1u;
  h ^= h >> 15;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  return h & (n_slots - 1);
}

int build_bytes_table(const std::vector<std::pair<std::string, int32_t>>& entries, SpTables* t) {
  uint32_t slots = 16;
  while (slots < entries.size() * 2 + 16) slots <<= 1;
  t->vocab_table.assign((size_t)slots * 4, 0u);
  t->vocab_blob.clear();
  for (const auto& e : entries) {
    if (e.first.empty()) continue;
    if (e.first.size() > 512) {
      t->error = "a vocabulary entry is longer than 512 bytes";
      return XLLM_ERR_UNSUPPORTED;
    }
    unsigned long long h = 0xcbf29ce484222325ull;
    for (unsigned char c : e.first) h = (h ^ c) * 0x100000001b3ull;
    if (h == 0) h = 1;
    if (t->vocab_blob.size() + e.first.size() >= (1u << 22)) {
      t->error = "vocabulary larger than 4 MiB of token bytes";
      return XLLM_ERR_UNSUPPORTED;
    }
    uint32_t slot = (uint32_t)(((h ^ (h >> 29)) * 0xBF58476D1CE4E5B9ull) >> 32) & (slots - 1);  // as hf_vocab_lookup
    while (t->vocab_table[(size_t)slot * 4] | t->vocab_table[(size_t)slot * 4 + 1]) slot = (slot + 1) & (slots - 1);
    t->vocab_table[(size_t)slot * 4 + 0] = (uint32_t)h;
    t->vocab_table[(size_t)slot * 4 + 1] = (uint32_t)(h >> 32);
    t->vocab_table[(size_t)slot * 4 + 2] = (uint32_t)e.second;
    t->vocab_table[(size_t)slot * 4 + 3] = ((uint32_t)t->vocab_blob.size() << 10) | (uint32_t)e.first.size();
    t->vocab_blob.insert(t->vocab_blob.end(), e.first.begin(), e.first.end());
  }
  return XLLM_OK;
}

int sp_load_model(const std::string& path_in, SpTables* t) {
  std::string path = path_in;
  struct stat st;
  if (stat(path.c_str(), &st) == 0 && S_ISDIR(st.st_mode)) path += "/tokenizer.model";  // tokenizer_args.h:37
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) {
    t->error = "cannot open " + path;
    return XLLM_ERR_IO;
  }
  std::string blob;
  char buf[1 << 16];
  size_t got;
  while ((got = fread(buf, 1, sizeof(buf), f)) > 0) blob.append(buf, got);
  fclose(f);

  RawModel m;
  if (!parse(blob, &m) || m.pieces.empty()) {
    t->error = path + ": not a SentencePiece ModelProto";
    return XLLM_ERR_FORMAT;
  }
  if (m.model_type != 2 && m.model_type != 1) {
    t->error = "SentencePiece model_type " + std::to_string(m.model_type) + ": only BPE (2) and UNIGRAM (1) are supported on device";
    return XLLM_ERR_UNSUPPORTED;
  }
  t->unigram = m.model_type == 1;
  if (!m.escape_ws || m.suffix_ws) {
    t->error = "escape_whitespaces=false / treat_whitespace_as_suffix=true are not supported on device";
    return XLLM_ERR_UNSUPPORTED;
  }
  const uint32_t P = (uint32_t)m.pieces.size();
  t->n_pieces = P;
  t->byte_fallback = m.byte_fallback;
  t->add_dummy_prefix = m.dummy_prefix;
  t->remove_extra_whitespaces = m.remove_ws;
  t->byte_id.assign(256, -1);
  t->piece_str.resize(P);
  t->piece_type.resize(P);

  // ---- InitializePieces (model_interface.cc): two maps, unk, byte pieces
  std::unordered_map<std::string_view, uint32_t> normal, reserved;
  int unk = -1;
  for (uint32_t i = 0; i < P; ++i) {
    const RawPiece& pc = m.pieces[i];
    t->piece_str[i] = pc.s;
    t->piece_type[i] = (uint8_t)pc.type;
    if (pc.s.empty()) { t->error = "piece must not be empty."; return XLLM_ERR_FORMAT; }
    if (pc.type == kUserDefined || pc.type == kUnused) {
      t->error = "USER_DEFINED / UNUSED pieces are not supported on device";
      return XLLM_ERR_UNSUPPORTED;
    }
    const bool is_normal = pc.type == kNormal;
    if (!(is_normal ? normal : reserved).emplace(std::string_view(m.pieces[i].s), i).second) {
      t->error = pc.s + " is already defined.";
      return XLLM_ERR_FORMAT;
    }
    if (pc.type == kUnknown) {
      if (unk >= 0) { t->error = "unk is already defined."; return XLLM_ERR_FORMAT; }
      unk = (int)i;
    }
    if (pc.type == kByte) {
      unsigned v = 0;
      if (!m.byte_fallback || pc.s.size() != 6 || sscanf(pc.s.c_str(), "<0x%02X>", &v) != 1) {
        t->error = "byte piece " + pc.s + " is invalid.";
        return XLLM_ERR_FORMAT;
      }
      t->byte_id[v & 0xFF] = (int32_t)i;
    }
  }
  if (unk < 0) { t->error = "unk is not defined."; return XLLM_ERR_FORMAT; }
  t->unk_id = unk;
  if (m.byte_fallback)
    for (int b = 0; b < 256; ++b)
      if (t->byte_id[b] < 0) { t->error = "there are not 256 byte pieces although `byte_fallback` is true."; return XLLM_ERR_FORMAT; }

  // ---- normalizer blob
  if (!m.charsmap.empty()) {
    uint32_t trie_bytes = 0;
    if (m.charsmap.size() <= 4) { t->error = "Blob for normalization rule is broken."; return XLLM_ERR_FORMAT; }
    memcpy(&trie_bytes, m.charsmap.data(), 4);
    if (trie_bytes >= m.charsmap.size() || trie_bytes % 4 != 0 || 4 + (size_t)trie_bytes > m.charsmap.size()) {
      t->error = "Trie data size exceeds the input blob size.";
      return XLLM_ERR_FORMAT;
    }
    t->trie.resize(trie_bytes / 4);
    memcpy(t->trie.data(), m.charsmap.data() + 4, trie_bytes);
    t->blob.assign(m.charsmap.begin() + 4 + trie_bytes, m.charsmap.end());
    t->blob.push_back(0);  // make every replacement NUL-terminated even if the blob is not
  }
  if (!t->trie.empty()) {
    // classify ASCII first bytes against the Darts trie (see SpTables::simple_ascii)
    const std::vector<uint32_t>& tr = t->trie;
    auto off = [](uint32_t u) { return (u >> 10) << ((u & 0x200u) >> 6); };
    const uint32_t root = off(tr[0]);
    for (uint32_t b = 0; b < 128; ++b) {
      bool simple = true;
      const uint32_t node = root ^ b;
      if (node < tr.size() && (tr[node] & 0x800000FFu) == b) {
        const uint32_t u = tr[node];
        if ((u >> 8) & 1u) simple = false;  // b alone is a key
        const uint32_t next = node ^ off(u);
        for (uint32_t c = 0; c < 128 && simple; ++c) {
          const uint32_t ch = next ^ c;
          if (ch < tr.size() && (tr[ch] & 0x800000FFu) == c && c != 0) simple = false;  // continues with ASCII
        }
      }
      if (b == 0) simple = false;  // NUL always takes the general path
      if (!simple) t->simple_ascii[b >> 5] &= ~(1u << (b & 31));
      // "space-like": b alone is a key whose replacement is exactly " ", and no longer key continues with ASCII —
      // then, followed by an ASCII byte, b normalises exactly like a source space (nmt_nfkc: \t \n \r ...)
      if (!simple && b != 0 && node < tr.size() && (tr[node] & 0x800000FFu) == b && ((tr[node] >> 8) & 1u)) {
        const uint32_t u = tr[node];
        const uint32_t next = node ^ off(u);
        bool longer_ascii = false;
        for (uint32_t c = 1; c < 128; ++c) {
          const uint32_t ch = next ^ c;
          if (ch < tr.size() && (tr[ch] & 0x800000FFu) == c) longer_ascii = true;
        }
        if (!longer_ascii && next < tr.size()) {
          const uint32_t val = tr[next] & 0x7FFFFFFFu;
          if ((size_t)val + 1 < t->blob.size() && t->blob[val] == ' ' && t->blob[val + 1] == 0)
            t->spacelike_ascii[b >> 5] |= 1u << (b & 31);
        }
      }
    }
  }
  {
    // the longest escaped replacement bounds what one unit can append (U+FFFD / identity: <= 4; ' ' -> 3)
    uint32_t cur = 0, mx = 4;
    for (uint8_t c : t->blob) {
      if (c == 0) { mx = std::max(mx, cur); cur = 0; }
      else cur += (c == ' ') ? 3 : 1;
    }
    t->max_unit_out = mx;
  }

  // ---- symbols: pieces + "virtual" single chars
  std::unordered_map<uint32_t, uint32_t> cp_sym;  // code point -> symbol
  std::vector<uint32_t> virt;                     // code points of virtual symbols
  auto sym_of_char = [&](const std::string& ch) -> uint32_t {
    uint32_t cp;
    if (!single_cp(ch, &cp)) return kEmptyKey;  // malformed char inside a piece: can never be matched
    auto it = cp_sym.find(cp);
    if (it != cp_sym.end()) return it->second;
    uint32_t sym;
    auto nit = normal.find(std::string_view(ch));
    if (nit != normal.end()) sym = nit->second;
    else { sym = P + (uint32_t)virt.size(); virt.push_back(cp); }
    cp_sym.emplace(cp, sym);
    return sym;
  };
  auto sym_of = [&](const std::string& s, size_t nchars) -> uint32_t {
    if (nchars == 1) return sym_of_char(s);
    auto it = normal.find(std::string_view(s));
    return it == normal.end() ? kEmptyKey : it->second;
  };
  // ranks: distinct scores, highest first
  std::vector<float> scores;
  for (const auto& pc : m.pieces)
    if (pc.type == kNormal) scores.push_back(pc.score);
  std::sort(scores.begin(), scores.end(), [](float a, float b) { return a > b; });
  scores.erase(std::unique(scores.begin(), scores.end()), scores.end());
  auto rank_of = [&](float s) {
    return (uint32_t)(std::lower_bound(scores.begin(), scores.end(), s, [](float a, float b) { return a > b; }) -
                      scores.begin());
  };

  // every char of every piece becomes a symbol (also single-char reserved pieces)
  const std::string kSpace = "\xe2\x96\x81";
  bool space_inside = false, space_only_after_space = true;
  std::vector<PairEntry> pairs;
  for (uint32_t i = 0; i < P; ++i) {
    const std::string& s = m.pieces[i].s;
    const auto b = char_bounds(s);
    const size_t nch = b.size() - 1;
    if (m.pieces[i].type != kNormal) {
      if (nch == 1) sym_of_char(s);
      continue;
    }
    for (size_t c = 0; c < nch; ++c) {
      const std::string ch = s.substr(b[c], b[c + 1] - b[c]);
      sym_of_char(ch);
      if (c > 0 && ch == kSpace) {
        space_inside = true;
        if (s.substr(b[c - 1], b[c] - b[c - 1]) != kSpace) space_only_after_space = false;
      }
    }
    if (nch < 2) continue;
    for (size_t c = 1; c < nch; ++c) {
      const std::string A = s.substr(0, b[c]), B = s.substr(b[c]);
      const uint32_t sa = sym_of(A, c), sb = sym_of(B, nch - c);
      if (sa == kEmptyKey || sb == kEmptyKey) continue;
      pairs.push_back(PairEntry{sa, sb, rank_of(m.pieces[i].score), i});
    }
  }
  t->space_sym = sym_of_char(kSpace);
  t->split_mode = !space_inside ? 1 : (space_only_after_space ? 2 : 0);
  if (t->unigram) {
    // unigram_model.cc: Viterbi over the NORMAL pieces; a word is scored on its own because no piece spans a
    // word start (split_mode), with the running float prefix score carried from word to word by the kernel
    if (t->split_mode == 0) {
      t->error = "Unigram model with a piece that holds U+2581 past its first char: no exact word split exists";
      return XLLM_ERR_UNSUPPORTED;
    }
    std::vector<std::pair<std::string, int32_t>> ent;
    float mn = 3.402823466e+38f;  // FLT_MAX, as upstream initialises min_score_
    t->piece_score.assign(P, 0.f);
    for (uint32_t i = 0; i < P; ++i) {
      t->piece_score[i] = m.pieces[i].score;
      if (m.pieces[i].type != kNormal) continue;
      ent.emplace_back(m.pieces[i].s, (int32_t)i);
      mn = m.pieces[i].score < mn ? m.pieces[i].score : mn;
      if (m.pieces[i].s.size() > t->max_piece_len) t->max_piece_len = (uint32_t)m.pieces[i].s.size();
    }
    if (t->max_piece_len > 64 || t->max_piece_len == 0) {
      t->error = "Unigram model without NORMAL pieces, or with a piece longer than 64 bytes";
      return XLLM_ERR_UNSUPPORTED;
    }
    t->unk_score = mn - 10.0f;  // min_score() - kUnkPenalty
    const int rc = build_bytes_table(ent, t);
    if (rc != XLLM_OK) return rc;
    pairs.clear();  // no merges in a Unigram model
    {
      // the piece trie upstream walks with Darts (unigram_model.cc trie_->traverse), as (parent, byte) -> child
      std::unordered_map<uint64_t, uint32_t> child;  // parent << 8 | byte -> node
      std::vector<int32_t> piece_of(1, -1);
      for (const auto& e : ent) {
        uint32_t node = 0;
        for (unsigned char c : e.first) {
          const uint64_t k = ((uint64_t)node << 8) | c;
          auto it = child.find(k);
          if (it == child.end()) {
            it = child.emplace(k, (uint32_t)piece_of.size()).first;
            piece_of.push_back(-1);
          }
          node = it->second;
        }
        piece_of[node] = e.second;
      }
      uint32_t slots = 16;
      while (slots < child.size() * 2 + 16) slots <<= 1;
      t->uni_trie.assign((size_t)slots * 4, 0xFFFFFFFFu);
      for (const auto& kv : child) {
        const uint32_t parent = (uint32_t)(kv.first >> 8), byte = (uint32_t)(kv.first & 0xFF);
        uint32_t slot = sp_trie_slot(parent, byte, slots);
        while (t->uni_trie[(size_t)slot * 4] != 0xFFFFFFFFu) slot = (slot + 1) & (slots - 1);
        t->uni_trie[(size_t)slot * 4 + 0] = parent;
        t->uni_trie[(size_t)slot * 4 + 1] = byte;
        t->uni_trie[(size_t)slot * 4 + 2] = kv.second;
        t->uni_trie[(size_t)slot * 4 + 3] = (uint32_t)piece_of[kv.second];
      }
    }
  }
  t->n_syms = P + (uint32_t)virt.size();
  t->virt_cp = virt;

  // ---- emission: PieceToId(symbol string) with the reserved map first
  t->emit.assign(t->n_syms, -1);
  for (uint32_t i = 0; i < P; ++i) {
    if (m.pieces[i].type != kNormal) continue;
    auto rit = reserved.find(std::string_view(m.pieces[i].s));
    uint32_t id = rit != reserved.end() ? rit->second : i;
    t->emit[i] = m.pieces[id].type == kUnknown ? -1 : (int32_t)id;
  }
  for (uint32_t v = 0; v < virt.size(); ++v) {
    // a virtual symbol is a single char that is not a NORMAL piece: reserved hit or unknown
    std::string ch;
    const uint32_t cp = virt[v];
    if (cp < 0x80) ch.push_back((char)cp);
    else if (cp < 0x800) { ch.push_back((char)(0xC0 | (cp >> 6))); ch.push_back((char)(0x80 | (cp & 0x3F))); }
    else if (cp < 0x10000) { ch.push_back((char)(0xE0 | (cp >> 12))); ch.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); ch.push_back((char)(0x80 | (cp & 0x3F))); }
    else { ch.push_back((char)(0xF0 | (cp >> 18))); ch.push_back((char)(0x80 | ((cp >> 12) & 0x3F))); ch.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); ch.push_back((char)(0x80 | (cp & 0x3F))); }
    auto rit = reserved.find(std::string_view(ch));
    if (rit != reserved.end() && m.pieces[rit->second].type != kUnknown) t->emit[P + v] = (int32_t)rit->second;
  }

  // ---- hash tables
  t->ascii_sym.assign(128, 0);
  for (uint32_t c = 0; c < 128; ++c) {
    auto it = cp_sym.find(c);
    t->ascii_sym[c] = it != cp_sym.end() ? it->second : (kSymUnknownFlag | c);
  }
  {
    const uint32_t n = pow2_at_least(cp_sym.size() * 4 + 16);
    t->cp_table.assign(n, CpEntry{kEmptyKey, 0});
    for (const auto& kv : cp_sym) {
      uint32_t h = sp_hash_cp(kv.first) & (n - 1);
      while (t->cp_table[h].cp != kEmptyKey) h = (h + 1) & (n - 1);
      t->cp_table[h] = CpEntry{kv.first, kv.second};
    }
  }
  {
    const uint32_t n = pow2_at_least(pairs.size() * 4 + 16);
    t->pair_table.assign(n, PairEntry{kEmptyKey, kEmptyKey, kNoPrio, 0});
    for (const auto& e : pairs) {
      uint32_t h = sp_pair_slot(e.a, e.b, n);
      while (t->pair_table[h].a != kEmptyKey) {
        if (t->pair_table[h].a == e.a && t->pair_table[h].b == e.b) break;  // cannot happen (A||B is unique)
        h = (h + 1) & (n - 1);
      }
      t->pair_table[h] = e;
    }
  }
  return XLLM_OK;
}

}  // namespace xllm
