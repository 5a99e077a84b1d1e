"""Natural text (source code and prose from this image's site-packages: punctuation, digits, newlines, indentation,
long identifiers, UTF-8) through tokenizers trained on that kind of text — SentencePiece BPE 32 000 and HF byte-level
BPE with a 128 471-entry vocabulary (ids beyond 16 bits) — frozen from pip sentencepiece / pip tokenizers by
tests/golden/make_natural_fixtures.py.  The CPU oracles must reproduce the wheels on all of it; the GPU side of the
same goldens is tests/test_gpu_natural.py."""
import json
import os

import pytest

HERE = os.path.dirname(__file__)
GOLD = os.path.join(HERE, "golden", "natural_goldens.json")


@pytest.fixture(scope="module")
def cases():
    with open(GOLD) as f:
        return json.load(f)["cases"]


def test_sentencepiece_oracle_on_natural_text(oracle, cases):
    S = oracle.SentencePieceOracle(os.path.join(HERE, "golden", "sp_natural_32k"))
    assert S.vocab_size() == 32000
    bad = [i for i, c in enumerate(cases) if S.encode(bytes.fromhex(c["text"])).tolist() != c["sp"]]
    assert not bad, bad[:5]


def test_hf_oracle_on_natural_text(oracle, cases):
    H = oracle.HfBpeOracle(os.path.join(HERE, "golden", "hf_natural_128k"))
    bad = [i for i, c in enumerate(cases)
           if H.prefix_ids + H.encode(bytes.fromhex(c["text"])).tolist() + H.suffix_ids != c["hf"]]
    assert not bad, bad[:5]
