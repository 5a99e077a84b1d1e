"""GPU parity of the HF `tokenizer.json` backend (csrc/hf_model.cc tables + the regex pre-tokenizer of
csrc/hf_pretok.cuh in front of the shared merge kernels) against the committed pip-`tokenizers` goldens and the
CPU oracle (fast_tokenizer.cpp:20-30 -> tokenizers_encode, lib.rs:83-99)."""
import ctypes
import json
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(__file__)
MODEL_DIR = os.path.join(HERE, "golden", "hf_bpe_8k")
GOLD = os.path.join(HERE, "golden", "hf_bpe_goldens.json")


@pytest.fixture(scope="module")
def tok():
    import xllm_service_b200 as x
    h = x.Ingest(tokenizer_path=MODEL_DIR)   # tokenizer.json present -> the HF backend
    yield h
    h.close()


def _encode_all(tok, texts):
    from xllm_service_b200 import workload
    b = workload.pack_prompts(texts)
    stride = max(16, max((len(t) for t in texts), default=0) + 8)
    ids, n_ids, status = tok.encode_batch(b.text, b.offsets, stride)
    return [ids[i, :n_ids[i]].tolist() for i in range(len(texts))], status


def test_goldens_and_vocab(tok):
    with open(GOLD) as f:
        g = json.load(f)
    assert tok.vocab_size() == g["vocab_size"]
    texts = [bytes.fromhex(c["text"]) for c in g["cases"]]
    got, status = _encode_all(tok, texts)
    assert (status == 0).all()
    bad = [(t[:40], a[:12], c["ids"][:12]) for t, a, c in zip(texts, got, g["cases"]) if a != c["ids"]]
    assert not bad, (len(bad), bad[:5])


ALPHABET = list("abcdefghij  \t\n'.,!?012") + [
    "é", "日", "Σ", "١", " ", "　", "\U0001F600", "'s", "'re", " '", "<|endoftext|>", "\r\n", " ", "_",
    "²", "", " ", " ", " ", " ", "ǅ", "Ⅷ", "́", "﻿", "​", "­", "᠎",
    "⁠", "'ll", "'t", "'", "''", "!'", " 's", "\n's", "'S", "<|endoftext", "|>"]


def test_fuzz_vs_oracle(tok, oracle):
    hf = oracle.HfBpeOracle(MODEL_DIR)
    rnd = random.Random(77)
    texts = []
    for _ in range(3000):
        texts.append("".join(rnd.choice(ALPHABET) for _ in range(rnd.randrange(0, 70))).encode())
    got, status = _encode_all(tok, texts)
    assert (status == 0).all()
    bad = [(t, a[:12], hf.encode(t).tolist()[:12]) for t, a in zip(texts, got) if a != hf.encode(t).tolist()]
    assert not bad, (len(bad), bad[:3])


def test_long_texts_cross_the_staging_buffer(tok, oracle):
    """Texts of 3 KB .. 60 KB: every decision near a buffer seam (look-ahead, contraction, added token, cut UTF-8
    char, the word list filling up) must come out as in one sequential scan."""
    hf = oracle.HfBpeOracle(MODEL_DIR)
    from xllm_service_b200 import workload
    rnd = random.Random(5)
    texts = []
    for n, seed in ((600, 1), (2500, 2), (9000, 3)):
        texts.append(" ".join(workload.sentences(1, (n, n), seed=seed)).encode())
    for k in range(60):
        texts.append("".join(rnd.choice(ALPHABET) for _ in range(rnd.randrange(800, 6000))).encode())
    texts.append(("a1" * 4000).encode())                  # 8000 one-byte pre-tokens: the word list caps out
    texts.append(("!a" * 3000 + "\n\n").encode())
    texts.append((" " * 5000 + "x").encode())             # one 4999-char whitespace token is too long ...
    texts.append(("日本語 " * 1500).encode())
    texts.append(("x <|endoftext|>" * 700).encode())
    texts.append(("don't we'll they're " * 500).encode())
    texts.append(("é" * 300 + " " + "z" * 700 + " " + "q" * 513).encode())   # 600 B / 700 B / 513 B pre-tokens
    for pad in range(1560, 1600):                          # slide an added token / contraction across the seam
        texts.append(("ab " * (pad // 3) + "x" * (pad % 3) + "<|endoftext|>'ll é　　y").encode())
    got, status = _encode_all(tok, texts)
    for i, (t, a) in enumerate(zip(texts, got)):
        want = hf.encode(t).tolist()
        if t.startswith(b" " * 5000):
            assert status[i] == -6                         # ... and is reported, not mis-tokenised
            continue
        assert status[i] == 0, (i, len(t), t[:30])
        assert a == want, (i, len(t), t[:30])


def test_malformed_utf8_is_an_error_per_request(tok):
    """The reference's Rust shim unwraps from_utf8 and panics (lib.rs:91); here the request fails, the batch
    does not."""
    texts = [b"fine", b"bad \xff byte", b"cut \xe6\x97", b"\x80start", b"over\xc0\xafong", b"sur\xed\xa0\x80r", b"ok too",
             b"ab " * 1000 + b"\xfe", b"\xf5\x80\x80\x80"]
    got, status = _encode_all(tok, texts)
    assert status.tolist() == [0, -1, -1, -1, -1, -1, 0, -1, -1]
    assert got[0] and got[6] and not got[1]


def test_template_post_processor_wraps_every_sequence(tmp_path, oracle):
    import xllm_service_b200 as x
    with open(os.path.join(MODEL_DIR, "tokenizer.json")) as f:
        d = json.load(f)
    d["post_processor"] = {
        "type": "TemplateProcessing",
        "single": [{"SpecialToken": {"id": "<|endoftext|>", "type_id": 0}}, {"Sequence": {"id": "A", "type_id": 0}},
                   {"SpecialToken": {"id": "<|endoftext|>", "type_id": 0}}],
        "pair": [{"Sequence": {"id": "A", "type_id": 0}}, {"Sequence": {"id": "B", "type_id": 1}}],
        "special_tokens": {"<|endoftext|>": {"id": "<|endoftext|>", "ids": [0], "tokens": ["<|endoftext|>"]}}}
    (tmp_path / "tokenizer.json").write_text(json.dumps(d))
    hf = oracle.HfBpeOracle(MODEL_DIR)
    texts = [b"", b"hello world", ("long " * 2000).encode()]
    h = x.Ingest(tokenizer_path=str(tmp_path))
    try:
        got, status = _encode_all(h, texts)
    finally:
        h.close()
    assert (status == 0).all()
    for t, a in zip(texts, got):
        assert a == [0] + hf.encode(t).tolist() + [0]
    try:
        from tokenizers import Tokenizer
    except ImportError:
        return
    ref = Tokenizer.from_file(str(tmp_path / "tokenizer.json"))
    for t, a in zip(texts, got):
        assert a == ref.encode(t.decode(), add_special_tokens=True).ids


def test_legacy_abi_round_trip(tok):
    """tokenizers_* C ABI (tokenizers.h:29-64) on the HF backend: encode, decode with and without special tokens,
    id <-> token in the vocabulary's byte-level spelling."""
    import xllm_service_b200 as x
    L = ctypes.CDLL(x.lib_path())
    L.tokenizers_new_from_path.restype = ctypes.c_void_p
    L.tokenizers_new_from_path.argtypes = [ctypes.c_char_p]
    class Result(ctypes.Structure):
        _fields_ = [("token_ids", ctypes.POINTER(ctypes.c_int)), ("len", ctypes.c_size_t)]
    L.tokenizers_encode.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int,
                                    ctypes.POINTER(Result)]
    L.tokenizers_free_encode_results.argtypes = [ctypes.POINTER(Result), ctypes.c_size_t]
    L.tokenizers_decode.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                    ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_size_t)]
    L.tokenizers_id_to_token.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.POINTER(ctypes.c_char_p),
                                         ctypes.POINTER(ctypes.c_size_t)]
    L.tokenizers_token_to_id.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t,
                                         ctypes.POINTER(ctypes.c_int32)]
    L.tokenizers_free.argtypes = [ctypes.c_void_p]
    h = L.tokenizers_new_from_path(MODEL_DIR.encode())
    assert h
    text = "byte level round trip é 日本<|endoftext|> tail".encode()
    res = Result()
    L.tokenizers_encode(h, text, len(text), 1, ctypes.byref(res))
    ids = [res.token_ids[i] for i in range(res.len)]
    L.tokenizers_free_encode_results(ctypes.byref(res), 1)
    assert ids == [int(v) for v in tok.encode(text)] and 0 in ids
    arr = (ctypes.c_uint32 * len(ids))(*ids)
    data, m = ctypes.c_char_p(), ctypes.c_size_t()
    L.tokenizers_decode(h, arr, len(ids), 0, ctypes.byref(data), ctypes.byref(m))
    assert ctypes.string_at(data, m.value) == text
    L.tokenizers_decode(h, arr, len(ids), 1, ctypes.byref(data), ctypes.byref(m))
    assert ctypes.string_at(data, m.value) == text.replace(b"<|endoftext|>", b"")
    with open(os.path.join(MODEL_DIR, "tokenizer.json")) as f:
        vocab = json.load(f)["model"]["vocab"]
    word, wid = next((k, v) for k, v in vocab.items() if k.startswith("Ġ") and len(k) > 3)
    L.tokenizers_id_to_token(h, wid, ctypes.byref(data), ctypes.byref(m))
    assert ctypes.string_at(data, m.value) == word.encode()
    out = ctypes.c_int32()
    L.tokenizers_token_to_id(h, word.encode(), len(word.encode()), ctypes.byref(out))
    assert out.value == wid
    L.tokenizers_free(h)
    # add_special_tokens = 0 leaves the template ids off (lib.rs:83-99); the Llama-3-style fixture has a BOS template
    h3 = L.tokenizers_new_from_path(os.path.join(HERE, "golden", "hf_llama3_style", "tokenizer.json").encode())
    assert h3
    got = {}
    for flag in (1, 0):
        res = Result()
        L.tokenizers_encode(h3, b"hello world", 11, flag, ctypes.byref(res))
        got[flag] = [res.token_ids[i] for i in range(res.len)]
        L.tokenizers_free_encode_results(ctypes.byref(res), 1)
    assert got[1][0] == 0 and got[1][1:] == got[0] and len(got[0]) >= 1
    L.tokenizers_free(h3)
