"""GPU parity of the tiktoken backend (csrc/tiktoken_model.cc tables through the same encode kernels) against
the committed pip-tiktoken goldens and the CPU oracle (tiktoken_tokenizer.cpp:115-294, regex-less mode)."""
import json
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(__file__)
MODEL_DIR = os.path.join(HERE, "golden", "tiktoken_1k")
GOLD = os.path.join(HERE, "golden", "tiktoken_goldens.json")


@pytest.fixture(scope="module")
def tok():
    import xllm_service_b200 as x
    h = x.Ingest(tokenizer_path=MODEL_DIR)   # tokenizer_config.json selects TikTokenTokenizer
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
    assert tok.vocab_size() == g["n_ranks"]
    texts = [bytes.fromhex(c["text"]) for c in g["cases"]]
    got, status = _encode_all(tok, texts)
    assert (status == 0).all()
    for t, a, c in zip(texts, got, g["cases"]):
        assert a == c["ids"], t[:40]


def test_fuzz_and_long_texts_vs_oracle(tok, oracle):
    tik = oracle.TiktokenOracle(MODEL_DIR)
    from xllm_service_b200 import workload
    rnd = random.Random(12)
    texts = [b"", b"a", b"\x00", b"hello\x00world\x7f\xf5!", b"\x00\x7f"]
    for _ in range(300):
        texts.append(bytes(rnd.choice(b"abcdefghijklmnop  \n\xc3\xa9\xe6\x97\xa5\x00xyz") for _ in range(rnd.randrange(1, 260))))
    # whole prompts are ONE piece in the reference's regex-less mode: 600 B .. 40 KB go through the
    # cooperative and the global-scratch paths
    for n, seed in ((30, 1), (120, 2), (600, 3), (2500, 4)):
        texts.append(" ".join(workload.sentences(1, (n, n), seed=seed)).encode())
    got, status = _encode_all(tok, texts)
    assert (status == 0).all()
    for t, a in zip(texts, got):
        assert a == tik.encode(t).tolist(), (len(t), t[:40])


def test_legacy_decode_round_trip(tok):
    import ctypes
    import xllm_service_b200 as x
    L = ctypes.CDLL(x.lib_path())
    L.tokenizers_new_from_path.restype = ctypes.c_void_p
    L.tokenizers_new_from_path.argtypes = [ctypes.c_char_p]
    L.tokenizers_decode.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                    ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_size_t)]
    L.tokenizers_free.argtypes = [ctypes.c_void_p]
    h = L.tokenizers_new_from_path(MODEL_DIR.encode())
    text = "byte level round trip é 日本".encode()
    ids = tok.encode(text)
    arr = (ctypes.c_uint32 * len(ids))(*ids)
    data, n = ctypes.c_char_p(), ctypes.c_size_t()
    L.tokenizers_decode(h, arr, len(ids), 0, ctypes.byref(data), ctypes.byref(n))
    assert ctypes.string_at(data, n.value) == text
    L.tokenizers_free(h)
