/*
 * tokenizers.h — the legacy tokenizer C-ABI, served by libxllm_ingest.so.
 *
 * These are the symbols the reference's only existing FFI seam declares
 * (xllm_service/tokenizer/tokenizers/tokenizers.h:29-64, implemented by the Rust shim
 * xllm_service/tokenizer/tokenizers/src/lib.rs:56-204 and consumed by
 * xllm_service/tokenizer/fast_tokenizer.cpp:8-78), so FastTokenizer links against this library
 * unchanged: same names, argument types, meaning and ownership.  Behaviour to know about:
 *   - the path given to tokenizers_new_from_path may be the `tokenizer.json` FastTokenizer passes (byte-level
 *     BPE in the GPT-2 or the Llama-3 / Qwen2 layout, csrc/hf_model.cc), or a tokenizer directory /
 *     `tokenizer.model` (SentencePiece BPE or Unigram, tiktoken); a model outside the supported envelope yields
 *     NULL with the reason in xllm_last_error() (FastTokenizer then CHECK-fails loudly, fast_tokenizer.cpp:10-11);
 *   - failures return NULL / leave the result empty instead of panicking (lib.rs:32,39,69,77,91);
 *   - encode runs on the GPU; there is no CPU fallback.
 * Handles are not thread-safe (the Rust ones are not either, lib.rs:139-154): one per thread.
 */
#ifndef XLLM_TOKENIZERS_H_
#define XLLM_TOKENIZERS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* TokenizerHandle;

/* ids of one encoded text; token_ids is allocated by the library (tokenizers_free_encode_results) */
typedef struct {
  int* token_ids;
  size_t len;
} TokenizerEncodeResult;

/* ---- lifetime: tokenizers.h:36,62 / lib.rs:65-80,157-161; from_str is exported by the shim (lib.rs:57-62) but
 * unused by the C++ side: the bytes of a serialized model */
TokenizerHandle tokenizers_new_from_path(const char* model_path);
TokenizerHandle tokenizers_new_from_str(const char* model_bytes, size_t n_bytes);
void tokenizers_free(TokenizerHandle tok);

/* ---- text -> ids: tokenizers.h:38-42 / lib.rs:83-136.  with_special_tokens = 0 leaves off the ids a
 * tokenizer.json's template wraps around the sequence (the reference's only caller passes 1,
 * fast_tokenizer.cpp:24); SentencePiece / tiktoken models add none either way
 * (sentencepiece_tokenizer.cpp:115-128). */
void tokenizers_encode(TokenizerHandle tok, const char* text, size_t text_len, int with_special_tokens,
                       TokenizerEncodeResult* out);
void tokenizers_encode_batch(TokenizerHandle tok, const char* const* texts, const size_t* text_lens, size_t n_texts,
                             int with_special_tokens, TokenizerEncodeResult* outs);
void tokenizers_free_encode_results(TokenizerEncodeResult* outs, size_t n_texts);

/* ---- ids -> text: tokenizers.h:44-49 / lib.rs:139-154.  *text_out points into the handle and stays valid until
 * the next call on it */
void tokenizers_decode(TokenizerHandle tok, const uint32_t* ids, size_t n_ids, int skip_special_tokens,
                       const char** text_out, size_t* text_len_out);

/* ---- vocabulary: tokenizers.h:51-60,64 / lib.rs:164-204.  token_to_id stores -1 when the token is unknown */
void tokenizers_id_to_token(TokenizerHandle tok, uint32_t id, const char** token_out, size_t* token_len_out);
void tokenizers_token_to_id(TokenizerHandle tok, const char* token, size_t token_len, int32_t* id_out);
void tokenizers_get_vocab_size(TokenizerHandle tok, size_t* size_out);

#ifdef __cplusplus
}
#endif
#endif /* XLLM_TOKENIZERS_H_ */
