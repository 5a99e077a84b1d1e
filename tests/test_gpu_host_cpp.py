"""The C++ host side above the C-ABI, compiled with g++ against include/*.h and run on the GPU box:
the micro-batcher (host/ingest_batcher.h) driven from 16 threads, and the legacy tokenizers_* ABI
(include/tokenizers.h) through ctypes.  Results are compared with the CPU oracle."""
import ctypes
import os
import struct
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(__file__)
ROOT = os.path.dirname(HERE)
MODEL_DIR = os.path.join(HERE, "golden", "sp_bpe_8k")


def test_micro_batcher_threads(oracle, tmp_path):
    from xllm_service_b200 import workload
    texts = [s.encode() for s in workload.sentences(300, (3, 120), seed=21)] + [b"", b"x", "é日".encode()]
    pf = tmp_path / "prompts.bin"
    with open(pf, "wb") as f:
        for t in texts:
            f.write(struct.pack("<I", len(t)) + t)
    exe = tmp_path / "batcher_main"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "xllm_service_b200", "host"),
                           os.path.join(HERE, "cpp", "batcher_main.cc"), "-o", str(exe),
                           "-L", os.path.join(ROOT, "xllm_service_b200"), "-lxllm_ingest",
                           "-Wl,-rpath," + os.path.join(ROOT, "xllm_service_b200")])
    p = subprocess.run([str(exe), MODEL_DIR, str(pf), "16", "64", "200"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    sp = oracle.SentencePieceOracle(MODEL_DIR)
    lines = p.stdout.strip().split("\n")
    assert len(lines) == len(texts)
    for t, line in zip(texts, lines):
        f = [int(v) for v in line.split()]
        assert f[0] == 0 and f[3] == len(f) - 4
        assert f[4:] == sp.encode(t).tolist()
        assert f[1] == 0 and f[2] == 1      # the only prefill-side / decode-side instances
    assert "batches=" in p.stderr
    n_batches = int(p.stderr.split("batches=")[1].split()[0])
    assert n_batches < len(texts)           # requests really were coalesced


def test_legacy_tokenizers_abi(oracle):
    import xllm_service_b200 as x
    L = ctypes.CDLL(x.lib_path())

    class Res(ctypes.Structure):
        _fields_ = [("token_ids", ctypes.POINTER(ctypes.c_int)), ("len", ctypes.c_size_t)]

    L.tokenizers_new_from_path.restype = ctypes.c_void_p
    L.tokenizers_new_from_path.argtypes = [ctypes.c_char_p]
    L.tokenizers_new_from_str.restype = ctypes.c_void_p
    L.tokenizers_new_from_str.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
    L.tokenizers_encode.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(Res)]
    L.tokenizers_encode_batch.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_char_p),
                                          ctypes.POINTER(ctypes.c_size_t), ctypes.c_size_t, ctypes.c_int,
                                          ctypes.POINTER(Res)]
    L.tokenizers_free_encode_results.argtypes = [ctypes.POINTER(Res), ctypes.c_size_t]
    L.tokenizers_decode.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                    ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_size_t)]
    L.tokenizers_id_to_token.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.POINTER(ctypes.c_char_p),
                                         ctypes.POINTER(ctypes.c_size_t)]
    L.tokenizers_token_to_id.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t,
                                         ctypes.POINTER(ctypes.c_int32)]
    L.tokenizers_get_vocab_size.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_size_t)]
    L.tokenizers_free.argtypes = [ctypes.c_void_p]
    sp = oracle.SentencePieceOracle(MODEL_DIR)

    assert L.tokenizers_new_from_path(b"/nonexistent/dir") is None
    h = L.tokenizers_new_from_path(MODEL_DIR.encode())
    assert h
    n = ctypes.c_size_t()
    L.tokenizers_get_vocab_size(h, ctypes.byref(n))
    assert n.value == 8000
    text = "hello world ﬁne 日本 \U0001F600 " * 30
    tb = text.encode()
    r = Res()
    L.tokenizers_encode(h, tb, len(tb), 1, ctypes.byref(r))
    ids = [r.token_ids[i] for i in range(r.len)]
    assert ids == sp.encode(tb).tolist()
    # decode round trip (SentencePiece Decode semantics; the normaliser maps the ligature to "fi")
    arr = (ctypes.c_uint32 * len(ids))(*ids)
    data, dlen = ctypes.c_char_p(), ctypes.c_size_t()
    L.tokenizers_decode(h, arr, len(ids), 1, ctypes.byref(data), ctypes.byref(dlen))
    dec = ctypes.string_at(data, dlen.value).decode()
    assert dec == " ".join(text.replace("ﬁ", "fi").split())
    L.tokenizers_free_encode_results(ctypes.byref(r), 1)
    # batch form
    texts = [b"alpha beta", b"", b"gamma", "é".encode() * 50]
    arr_p = (ctypes.c_char_p * len(texts))(*texts)
    arr_l = (ctypes.c_size_t * len(texts))(*[len(t) for t in texts])
    res = (Res * len(texts))()
    L.tokenizers_encode_batch(h, arr_p, arr_l, len(texts), 0, res)
    for t, rr in zip(texts, res):
        assert [rr.token_ids[i] for i in range(rr.len)] == sp.encode(t).tolist()
    L.tokenizers_free_encode_results(res, len(texts))
    # vocabulary queries
    tid = ctypes.c_int32()
    L.tokenizers_token_to_id(h, "▁y".encode(), len("▁y".encode()), ctypes.byref(tid))
    assert tid.value == 259
    L.tokenizers_token_to_id(h, b"definitely-not-a-piece", 22, ctypes.byref(tid))
    assert tid.value == -1
    L.tokenizers_id_to_token(h, 259, ctypes.byref(data), ctypes.byref(dlen))
    assert ctypes.string_at(data, dlen.value) == "▁y".encode()
    # from the serialized model bytes
    blob = open(os.path.join(MODEL_DIR, "tokenizer.model"), "rb").read()
    h2 = L.tokenizers_new_from_str(blob, len(blob))
    assert h2
    L.tokenizers_encode(h2, b"hello", 5, 0, ctypes.byref(r))
    assert [r.token_ids[i] for i in range(r.len)] == sp.encode(b"hello").tolist()
    L.tokenizers_free_encode_results(ctypes.byref(r), 1)
    L.tokenizers_free(h2)
    L.tokenizers_free(h)
