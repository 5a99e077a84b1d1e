/*
 * tokenizers.h — the legacy tokenizer C-ABI, served by libxllm_ingest.so.
 *
 * These are the symbols the reference's only existing FFI seam declares
 * (xllm_service/tokenizer/tokenizers/tokenizers.h:29-64, implemented by the Rust shim
 * xllm_service/tokenizer/tokenizers/src/lib.rs:56-204 and consumed by
 * xllm_service/tokenizer/fast_tokenizer.cpp:8-78), so FastTokenizer links against this library
 * unchanged.  Same names, argument meaning and ownership; differences in behaviour:
 *   - the model behind a handle is the SentencePiece-BPE `tokenizer.model` of the directory (or
 *     file) passed to tokenizers_new_from_path; a HF `tokenizer.json` is not accepted yet and
 *     yields NULL (FastTokenizer then CHECK-fails loudly, fast_tokenizer.cpp:10-11);
 *   - failures return NULL / leave *result empty instead of panicking (lib.rs:32,39,69,77,91);
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

typedef struct {
  int* token_ids; /* callee-allocated; release with tokenizers_free_encode_results */
  size_t len;
} TokenizerEncodeResult;

/* tokenizers.h:36 / lib.rs:65-80 */
TokenizerHandle tokenizers_new_from_path(const char* path);
/* lib.rs:57-62 (exported by the shim, unused by the C++ side): the bytes of a serialized model */
TokenizerHandle tokenizers_new_from_str(const char* data, size_t len);
/* tokenizers.h:38-42 / lib.rs:83-99.  add_special_token = 0 leaves off the ids a tokenizer.json's template wraps
 * around the sequence (the reference's only caller passes 1, fast_tokenizer.cpp:24); SentencePiece / tiktoken
 * models add none either way (sentencepiece_tokenizer.cpp:115-128). */
void tokenizers_encode(TokenizerHandle handle, const char* data, size_t len, int add_special_token,
                       TokenizerEncodeResult* result);
/* lib.rs:102-126 */
void tokenizers_encode_batch(TokenizerHandle handle, const char* const* data, const size_t* len, size_t num_seqs,
                             int add_special_token, TokenizerEncodeResult* results);
/* lib.rs:129-136 */
void tokenizers_free_encode_results(TokenizerEncodeResult* results, size_t num_seqs);
/* tokenizers.h:44-49 / lib.rs:139-154: *decode_data points into the handle, valid until the next call */
void tokenizers_decode(TokenizerHandle handle, const uint32_t* data, size_t len, int skip_special_tokens,
                       const char** decode_data, size_t* decode_len);
/* tokenizers.h:51-54 / lib.rs:171-187 */
void tokenizers_id_to_token(TokenizerHandle handle, uint32_t id, const char** data, size_t* len);
/* tokenizers.h:56-60 / lib.rs:190-204: stores -1 to *id if the token is not in the vocab */
void tokenizers_token_to_id(TokenizerHandle handle, const char* token, size_t len, int32_t* id);
/* tokenizers.h:62 / lib.rs:157-161 */
void tokenizers_free(TokenizerHandle handle);
/* tokenizers.h:64 / lib.rs:164-168 */
void tokenizers_get_vocab_size(TokenizerHandle handle, size_t* size);

#ifdef __cplusplus
}
#endif
#endif /* XLLM_TOKENIZERS_H_ */
