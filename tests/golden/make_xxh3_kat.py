#!/usr/bin/env python3
"""Mint tests/golden/xxh3_kat.json from the real upstream xxHash (libxxhash.so.0.8.2, the
library the reference links as third_party/xxHash).  The reference itself holds no
known-answer vectors for this path, so these are minted here.

Run:  python tests/golden/make_xxh3_kat.py     (needs libxxhash.so.0; this container has it)
"""
import ctypes
import json
import os
import random
import struct

x = ctypes.CDLL("libxxhash.so.0")


class H(ctypes.Structure):
    _fields_ = [("lo", ctypes.c_uint64), ("hi", ctypes.c_uint64)]


x.XXH3_128bits_withSeed.restype = H
x.XXH3_128bits_withSeed.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_uint64]
x.XXH_versionNumber.restype = ctypes.c_uint


def h128(data, seed):
    h = x.XXH3_128bits_withSeed(data, len(data), seed)
    return struct.pack("<QQ", h.lo, h.hi)  # == memcpy(&XXH128_hash_t) on LE (hash_util.cpp:26-27)


def chain(tokens, block_size, seed):
    keys = []
    prev = None
    for b in range(len(tokens) // block_size):
        blk = struct.pack("<%di" % block_size, *tokens[b * block_size:(b + 1) * block_size])
        prev = h128(blk if prev is None else prev + blk, seed)
        keys.append(prev.hex())
    return keys


def main():
    rnd = random.Random(20260921)
    out = {"xxhash_version": x.XXH_versionNumber(), "raw": [], "chains": []}
    # raw XXH3_128bits_withSeed vectors across all four length classes + the >1024 scramble path
    for n in [0, 1, 2, 3, 4, 7, 8, 9, 15, 16, 17, 31, 32, 33, 64, 65, 96, 97, 127, 128, 129, 160, 161, 192, 239, 240,
              241, 255, 256, 272, 511, 512, 513, 527, 528, 529, 1019, 1020, 1023, 1024, 1025, 2048, 3000]:
        for seed in (0, 1024, 0x9E3779B185EBCA87):
            data = bytes(rnd.getrandbits(8) for _ in range(n))
            out["raw"].append({"len": n, "seed": seed, "data": data.hex(), "hash": h128(data, seed).hex()})
    # chain vectors: (tokens spec, block_size, seed)
    specs = [("iota256", list(range(256)), 128, 1024)]
    for (name, n, bs, seed) in [("rand_4096_bs128", 4096, 128, 1024), ("rand_300_bs128", 300, 128, 1024),
                                ("rand_127_bs128", 127, 128, 1024), ("rand_1000_bs16", 1000, 16, 1024),
                                ("rand_999_bs251", 999, 251, 1024), ("rand_64_bs1", 64, 1, 1024),
                                ("rand_200_bs3", 200, 3, 7), ("rand_400_bs30", 400, 30, 0),
                                ("rand_2048_bs128_seed0", 2048, 128, 0), ("rand_640_bs56", 640, 56, 1024),
                                ("rand_512_bs57", 512, 57, 123456789)]:
        specs.append((name, [rnd.randrange(-2**31, 2**31) if i % 7 == 0 else rnd.randrange(0, 152000)
                             for i in range(n)], bs, seed))
    for name, toks, bs, seed in specs:
        out["chains"].append({"name": name, "block_size": bs, "seed": seed, "tokens": toks,
                              "keys": chain(toks, bs, seed)})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "xxh3_kat.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes; iota256 keys:", out["chains"][0]["keys"])


if __name__ == "__main__":
    main()
