"""Pins the CPU oracle's XXH3-128 chain (oracle/xxh3_oracle.c) against
(1) the committed known-answer vectors minted from upstream libxxhash 0.8.2
    (tests/golden/xxh3_kat.json, generator: tests/golden/make_xxh3_kat.py),
(2) the SURVEY.md §8(c) KAT, and
(3) libxxhash.so.0 live, when it is installed on this machine.
Reference semantics: xllm_service/common/hash_util.cpp:18-45,
xllm_service/scheduler/managers/global_kvcache_mgr.cpp:76-94."""
import ctypes
import json
import os
import random
import struct

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "xxh3_kat.json")


@pytest.fixture(scope="module")
def kat():
    with open(GOLD) as f:
        return json.load(f)


def test_raw_vectors(oracle, kat):
    assert len(kat["raw"]) > 100
    for v in kat["raw"]:
        got = oracle.xxh3_128_with_seed(bytes.fromhex(v["data"]), v["seed"]).hex()
        assert got == v["hash"], (v["len"], v["seed"])


def test_chain_vectors(oracle, kat):
    for c in kat["chains"]:
        keys = oracle.block_hash_chain(c["tokens"], c["block_size"], c["seed"])
        assert [bytes(k).hex() for k in keys] == c["keys"], c["name"]


def test_survey_kat(oracle):
    # int32 LE 0..255, block 128, seed 1024
    keys = oracle.block_hash_chain(np.arange(256, dtype=np.int32), 128, 1024)
    assert bytes(keys[0]).hex() == "a6c0e2fc92c32b1c1ccff29b710ca0d2"
    assert bytes(keys[1]).hex() == "f981264b143b2d5f81fd86ed5d58d2a1"
    # layout check: struct bytes are the reverse of the canonical (big-endian) digest
    h = oracle.xxh3_128_with_seed(struct.pack("<16i", *range(16)), 1024)
    assert h.hex() == "1247ae96b541bccd1d2bea48ea0a97e4"


def test_single_call_semantics(oracle):
    t = np.arange(100, 228, dtype=np.int32)
    k0 = oracle.xxh3_128bits_hash(None, t)
    k1 = oracle.xxh3_128bits_hash(k0, t)
    assert k0 != k1
    # chained == unchained hash of prev || tokens
    assert k1 == oracle.xxh3_128_with_seed(k0 + t.tobytes(), 1024)
    # hash_util.cpp:33 CHECK_GT(1024, 16 + 4n): n = 252 fails, n = 251 passes
    oracle.xxh3_128bits_hash(k0, np.zeros(251, dtype=np.int32))
    with pytest.raises(ValueError):
        oracle.xxh3_128bits_hash(k0, np.zeros(252, dtype=np.int32))
    # the unchained branch has no such limit
    oracle.xxh3_128bits_hash(None, np.zeros(5000, dtype=np.int32))


def test_tail_tokens_ignored_and_empty(oracle):
    t = np.arange(300, dtype=np.int32)
    assert oracle.block_hash_chain(t, 128).shape == (2, 16)
    assert oracle.block_hash_chain(t[:127], 128).shape == (0, 16)
    assert oracle.block_hash_chain(t[:0], 128).shape == (0, 16)
    assert (oracle.block_hash_chain(t, 128) == oracle.block_hash_chain(t[:256], 128)).all()


def test_live_against_libxxhash(oracle):
    try:
        x = ctypes.CDLL("libxxhash.so.0")
    except OSError:
        pytest.skip("libxxhash.so.0 not installed")

    class H(ctypes.Structure):
        _fields_ = [("lo", ctypes.c_uint64), ("hi", ctypes.c_uint64)]

    x.XXH3_128bits_withSeed.restype = H
    x.XXH3_128bits_withSeed.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_uint64]
    rnd = random.Random(7)
    for n in list(range(0, 260)) + [511, 512, 528, 1020, 1024, 1025, 4097, 9999]:
        for seed in (0, 1024, rnd.getrandbits(64)):
            d = bytes(rnd.getrandbits(8) for _ in range(n))
            h = x.XXH3_128bits_withSeed(d, n, seed)
            assert oracle.xxh3_128_with_seed(d, seed) == struct.pack("<QQ", h.lo, h.hi), (n, seed)
