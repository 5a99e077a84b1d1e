"""GPU parity of the DeepSeek-V3 / R1 `tokenizer.json` layout — Sequence[Split(\\p{N}{1,3}), Split(CJK + kana runs),
Split(main regex), ByteLevel(use_regex=false)], all Isolated, normalizer = empty Sequence (csrc/hf_model.cc +
hf_pretok.cuh pattern 3) — against pip-`tokenizers` goldens (tests/golden/make_hf_fixture3.py) and the CPU oracle
(fast_tokenizer.cpp:20-30 -> tokenizers_encode; the family is named at scheduler/xllm_chat_parse_bridge.cpp:49-78)."""
import json
import os
import random
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(__file__)
DIR = os.path.join(HERE, "golden", "hf_deepseek_style")
GOLD = os.path.join(HERE, "golden", "hf_deepseek_goldens.json")
sys.path.insert(0, os.path.join(HERE, "golden"))


def _encode_all(tok, texts):
    from xllm_service_b200 import workload
    b = workload.pack_prompts(texts)
    stride = max(16, max((len(t) for t in texts), default=0) + 8)
    ids, n_ids, status = tok.encode_batch(b.text, b.offsets, stride)
    return [ids[i, :n_ids[i]].tolist() for i in range(len(texts))], status


@pytest.fixture(scope="module")
def setup(oracle):
    import xllm_service_b200 as x
    h = x.Ingest(tokenizer_path=DIR)
    yield h, oracle.HfBpeOracle(DIR)
    h.close()


def test_goldens(setup):
    tok, _ = setup
    with open(GOLD) as f:
        cases = json.load(f)["cases"]
    texts = [bytes.fromhex(c["text"]) for c in cases]
    got, status = _encode_all(tok, texts)
    assert (status == 0).all(), [(texts[i][:30], int(status[i])) for i in np.nonzero(status)[0][:5]]
    bad = [(t[:40], a[:12], c["ids"][:12]) for t, a, c in zip(texts, got, cases) if a != c["ids"]]
    assert not bad, (len(bad), bad[:5])


def test_fuzz_vs_oracle(setup):
    import make_hf_fixture3 as m
    tok, hf = setup
    rnd = random.Random(17)
    texts = ["".join(rnd.choice(m.ALPHABET) for _ in range(rnd.randrange(0, 70))).encode() for _ in range(6000)]
    got, status = _encode_all(tok, texts)
    assert (status == 0).all()
    bad = [(t, a[:12], hf.encode(t).tolist()[:12]) for t, a in zip(texts, got) if a != hf.encode(t).tolist()]
    assert not bad, (len(bad), bad[:3])


def test_long_texts_cross_the_staging_buffer(setup):
    """3 KB .. 40 KB: the carries of the four scans (swallowed newline, digit index, newline ahead, ASCII letters of a
    "punct + letters" token), whitespace runs touching the end of a buffer, prefixes, CJK runs, number groups and
    added tokens sliding across the seam."""
    import make_hf_fixture3 as m
    from xllm_service_b200 import workload
    tok, hf = setup
    rnd = random.Random(8)
    texts = []
    for n, seed in ((600, 1), (2500, 2)):
        texts.append(" ".join(workload.sentences(1, (n, n), seed=seed)).encode())
    for k in range(80):
        texts.append("".join(rnd.choice(m.ALPHABET) for _ in range(rnd.randrange(800, 5000))).encode())
    texts.append(("1234567" * 900).encode())
    texts.append(("a1" * 4000).encode())
    texts.append(("x \n \n  y!\n\n" * 500).encode())
    texts.append(("  \n" * 30 + "z" + " " * 300 + "\n" + " " * 299 + "q").encode())
    texts.append((".com@user#tag !x.y " * 500).encode())
    texts.append((("." + "abcdefghij" * 90 + "é ") * 8).encode())  # 900-letter "punct + letters" tokens, then a break
    big = ("." + "abcdefghij" * 700 + "é").encode()                # one 7 KB pre-token: over the ~1 KB device scratch
    texts.append(("日本語テキスト123 ひらがな " * 700).encode())
    texts.append((("日" * 300 + "ab") * 30).encode())               # 900-byte CJK runs (one token each) across buffers
    texts.append(("x <|endoftext|>\n" * 600).encode())
    texts.append(("\x01\x02a \x01 b\x7f\x7fc " * 700).encode())
    for pad in range(1550, 1600):
        texts.append(("ab " * (pad // 3) + "x" * (pad % 3) + " \n \n!ab 12345\t!a<|endoftext|> é　　y日本.z\x01q").encode())
    got, status = _encode_all(tok, texts + [big])
    assert status[-1] == -6                                        # refused loudly (XLLM_ERR_CAPACITY), never split wrongly
    for i, (t, a) in enumerate(zip(texts, got)):
        assert status[i] == 0, (i, len(t), t[:30])
        assert a == hf.encode(t).tolist(), (i, len(t), t[:30])
