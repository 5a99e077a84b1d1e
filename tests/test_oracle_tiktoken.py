"""Pins the CPU oracle of the reference's tiktoken backend (oracle/tiktoken_oracle.cc, restating
xllm_service/tokenizer/tiktoken_tokenizer.cpp:115-294 in the regex-less mode the service runs) against
upstream pip tiktoken: committed goldens + live when the wheel is importable."""
import base64
import json
import os
import random

import pytest

HERE = os.path.dirname(__file__)
MODEL_DIR = os.path.join(HERE, "golden", "tiktoken_1k")
GOLD = os.path.join(HERE, "golden", "tiktoken_goldens.json")


@pytest.fixture(scope="module")
def tik(oracle):
    return oracle.TiktokenOracle(MODEL_DIR)


def test_goldens(tik):
    with open(GOLD) as f:
        g = json.load(f)
    assert tik.vocab_size() == g["n_ranks"]
    assert len(g["cases"]) > 60
    for c in g["cases"]:
        t = bytes.fromhex(c["text"])
        assert tik.encode(t).tolist() == c["ids"], t[:40]


def test_parts_without_a_rank_are_skipped(tik):
    # tiktoken_tokenizer.cpp:222-233: a part that is not in the encoder is logged and skipped; the fixture lacks
    # the single bytes 0x00, 0x7F, 0xF5, which can never merge with anything
    ab = tik.encode(b"hello world").tolist()
    assert tik.encode(b"hello world\x00").tolist() == ab
    assert tik.encode(b"\x7fhello world").tolist() == ab
    assert tik.encode(b"hello\xf5 world").tolist() == tik.encode(b"hello").tolist() + tik.encode(b" world").tolist()
    assert tik.encode(b"").size == 0


def test_live_against_tiktoken_wheel(tik):
    tiktoken = pytest.importorskip("tiktoken")
    ranks = {}
    for line in open(os.path.join(MODEL_DIR, "tokenizer.model")):
        tok, r = line.split()
        ranks[base64.b64decode(tok)] = int(r)
    enc = tiktoken.Encoding("fixture", pat_str=r"[\\s\\S]+", mergeable_ranks=ranks, special_tokens={})
    rnd = random.Random(3)
    for _ in range(400):
        t = bytes(rnd.choice(b"abcdefghijklmnop  \\n\\xc3\\xa9\\xe6\\x97\\xa5xyz") for _ in range(rnd.randrange(1, 200)))
        assert tik.encode(t).tolist() == enc._encode_single_piece(t), t
