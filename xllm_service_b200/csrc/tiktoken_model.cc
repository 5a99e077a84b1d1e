// tiktoken_model.cc — tiktoken vocabulary file -> the same device tables the SentencePiece path uses.
//
// Host-side replacement of TiktokenTokenizer's constructor (xllm_service/tokenizer/tiktoken_tokenizer.cpp:38-45,
// load_vocab :115-153: one `base64(token) SP rank` per line).  The service never configures a regex
// pattern or special tokens (tokenizer_args.cpp:30-71), so encode == byte_pair_encode over the WHOLE text
// (:236-241): symbols are single bytes, a pair (A, B) may merge iff the bytes A||B have a rank, lower rank
// first, leftmost on ties (:197-207); parts without a rank are skipped at the end (:222-233).
// In table form: byte_mode (every byte is a symbol, no normaliser, no whitespace rules), split_mode 0
// (no pre-split: the whole prompt is one pre-token and goes through the long-word path when it is long).
#include <stdio.h>
#include <string.h>
#include <sys/stat.h>

#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "sp_model.h"

namespace xllm {

uint32_t sp_pair_slot(uint32_t a, uint32_t b, uint32_t n_slots);

namespace {

bool b64_decode(std::string_view in, std::string* out) {
  int8_t T[256];
  memset(T, -1, sizeof(T));
  const char* a = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
  for (int i = 0; i < 64; ++i) T[(uint8_t)a[i]] = (int8_t)i;
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

}  // namespace

// True when <dir>/tokenizer_config.json names the tiktoken backend
// (tokenizer_factory.cpp:20-25: tokenizer_class == "TikTokenTokenizer").
bool tokenizer_dir_is_tiktoken(const std::string& dir) {
  std::string cls;
  return json_file_top_level_string(dir + "/tokenizer_config.json", "tokenizer_class", &cls) &&
         cls == "TikTokenTokenizer";
}

int tiktoken_load_model(const std::string& path_in, SpTables* t) {
  std::string path = path_in;
  struct stat st;
  if (stat(path.c_str(), &st) == 0 && S_ISDIR(st.st_mode)) path += "/tokenizer.model";  // tokenizer_args.h:37
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) {
    t->error = "Failed to open vocab file: " + path;
    return XLLM_ERR_IO;
  }
  std::vector<std::string> tokens;   // symbol -> bytes
  std::vector<int32_t> ranks;        // symbol -> rank (= token id)
  std::unordered_map<std::string, uint32_t> sym_of;
  char line[1 << 16];
  while (fgets(line, sizeof(line), f)) {
    std::string l(line);
    while (!l.empty() && (l.back() == '\n' || l.back() == '\r')) l.pop_back();
    if (l.empty()) continue;
    const size_t sp = l.find(' ');
    if (sp == std::string::npos || l.find(' ', sp + 1) != std::string::npos) continue;  // "Failed to parse line"
    std::string tok;
    if (!b64_decode(std::string_view(l).substr(0, sp), &tok) || tok.empty()) continue;
    char* endp = nullptr;
    const long rank = strtol(l.c_str() + sp + 1, &endp, 10);
    if (endp == l.c_str() + sp + 1 || *endp != '\0' || rank < 0 || rank >= 0x3FFFFFFF) continue;
    if (!sym_of.emplace(tok, (uint32_t)tokens.size()).second) continue;  // "Duplicate token"
    tokens.push_back(tok);
    ranks.push_back((int32_t)rank);
  }
  fclose(f);
  if (tokens.empty()) {
    t->error = path + ": no `base64 rank` lines";
    return XLLM_ERR_FORMAT;
  }
  const uint32_t V = (uint32_t)tokens.size();
  t->byte_mode = true;
  t->split_mode = 0;
  t->add_dummy_prefix = false;
  t->remove_extra_whitespaces = false;
  t->byte_fallback = false;
  t->unk_id = -1;
  t->max_unit_out = 4;
  t->n_pieces = V;
  // single bytes without a rank become virtual symbols that emit nothing
  t->ascii_sym.assign(256, 0);
  uint32_t n_syms = V;
  for (uint32_t b = 0; b < 256; ++b) {
    auto it = sym_of.find(std::string(1, (char)b));
    t->ascii_sym[b] = it != sym_of.end() ? it->second : n_syms++;
  }
  t->n_syms = n_syms;
  t->emit.assign(n_syms, -2);  // -2: skipped at emission (tiktoken_tokenizer.cpp:228-229)
  for (uint32_t s = 0; s < V; ++s) t->emit[s] = ranks[s];
  t->virt_cp.assign(n_syms - V + 1, 0);
  t->byte_id.assign(256, -1);
  t->cp_table.assign(16, CpEntry{kEmptyKey, 0});
  t->space_sym = kEmptyKey;
  // decode tables: piece_str indexed by token id where dense enough
  int32_t max_rank = 0;
  for (int32_t r : ranks) max_rank = r > max_rank ? r : max_rank;
  if ((uint32_t)max_rank < 4u * V + 1024u) {
    t->piece_str.assign((size_t)max_rank + 1, std::string());
    t->piece_type.assign((size_t)max_rank + 1, 1);
    for (uint32_t s = 0; s < V; ++s) t->piece_str[(size_t)ranks[s]] = tokens[s];
  }
  t->vocab_size_override = (int32_t)V;
  // pairs: every split of every multi-byte token whose halves are symbols
  std::vector<PairEntry> pairs;
  for (uint32_t s = 0; s < V; ++s) {
    const std::string& tk = tokens[s];
    for (size_t k = 1; k < tk.size(); ++k) {
      const std::string A = tk.substr(0, k), B = tk.substr(k);
      uint32_t sa, sb;
      if (A.size() == 1) sa = t->ascii_sym[(uint8_t)A[0]];
      else { auto it = sym_of.find(A); if (it == sym_of.end()) continue; sa = it->second; }
      if (B.size() == 1) sb = t->ascii_sym[(uint8_t)B[0]];
      else { auto it = sym_of.find(B); if (it == sym_of.end()) continue; sb = it->second; }
      pairs.push_back(PairEntry{sa, sb, (uint32_t)ranks[s], s});
    }
  }
  uint32_t n = 16;
  while (n < pairs.size() * 4 + 16) n <<= 1;
  t->pair_table.assign(n, PairEntry{kEmptyKey, kEmptyKey, kNoPrio, 0});
  for (const auto& e : pairs) {
    uint32_t h = sp_pair_slot(e.a, e.b, n);
    while (t->pair_table[h].a != kEmptyKey) h = (h + 1) & (n - 1);
    t->pair_table[h] = e;
  }
  return XLLM_OK;
}

int load_tokenizer_tables(const std::string& path, SpTables* out) {
  // TokenizerFactory::create_tokenizer (tokenizer_factory.cpp:9-32): tokenizer.json -> HF fast tokenizer
  // (hf_model.cc), tokenizer_class TikTokenTokenizer -> tiktoken, otherwise SentencePiece
  struct stat st;
  const bool is_dir = stat(path.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
  if (is_dir && tokenizer_dir_has_hf_json(path)) return hf_load_model(path, out);
  if (!is_dir && path.size() > 5 && path.compare(path.size() - 5, 5, ".json") == 0) return hf_load_model(path, out);
  if (is_dir && tokenizer_dir_is_tiktoken(path)) return tiktoken_load_model(path, out);
  return sp_load_model(path, out);
}

}  // namespace xllm
