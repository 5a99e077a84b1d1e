"""GPU parity of the newer HF `tokenizer.json` layouts (Llama-3 style: Split(cl100k-family regex, \\p{N}{1,3}) +
ByteLevel(use_regex=false), ignore_merges, BOS template; Qwen2 style: \\p{N}, normalizer NFC) — csrc/hf_model.cc +
hf_pretok.cuh pattern 2 — against pip-`tokenizers` goldens (tests/golden/make_hf_fixture2.py) and the CPU oracle
(fast_tokenizer.cpp:20-30 -> tokenizers_encode, lib.rs:83-99)."""
import json
import os
import random
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(__file__)
GOLD = os.path.join(HERE, "golden", "hf_cl100k_goldens.json")
STYLES = ["hf_llama3_style", "hf_qwen2_style"]
sys.path.insert(0, os.path.join(HERE, "golden"))


def _encode_all(tok, texts):
    from xllm_service_b200 import workload
    b = workload.pack_prompts(texts)
    stride = max(16, max((len(t) for t in texts), default=0) + 8)
    ids, n_ids, status = tok.encode_batch(b.text, b.offsets, stride)
    return [ids[i, :n_ids[i]].tolist() for i in range(len(texts))], status


@pytest.fixture(scope="module", params=STYLES)
def setup(request, oracle):
    import xllm_service_b200 as x
    d = os.path.join(HERE, "golden", request.param)
    h = x.Ingest(tokenizer_path=d)
    yield request.param, h, oracle.HfBpeOracle(d)
    h.close()


def test_goldens(setup):
    style, tok, _ = setup
    with open(GOLD) as f:
        cases = json.load(f)["cases"][style]
    texts = [bytes.fromhex(c["text"]) for c in cases]
    got, status = _encode_all(tok, texts)
    assert (status == 0).all(), [(texts[i][:30], int(status[i])) for i in np.nonzero(status)[0][:5]]
    bad = [(t[:40], a[:12], c["ids"][:12]) for t, a, c in zip(texts, got, cases) if a != c["ids"]]
    assert not bad, (len(bad), bad[:5])


def test_fuzz_vs_oracle(setup):
    import make_hf_fixture2 as m
    _, tok, hf = setup
    rnd = random.Random(99)
    texts = ["".join(rnd.choice(m.ALPHABET) for _ in range(rnd.randrange(0, 70))).encode() for _ in range(4000)]
    got, status = _encode_all(tok, texts)
    assert (status == 0).all()
    bad = []
    for t, a in zip(texts, got):
        want = hf.prefix_ids + hf.encode(t).tolist() + hf.suffix_ids
        if a != want:
            bad.append((t, a[:12], want[:12]))
    assert not bad, (len(bad), bad[:3])


def test_long_texts_cross_the_staging_buffer(setup):
    """3 KB .. 40 KB: the carries of the swallowed-newline / digit-index / newline-ahead scans, whitespace runs that
    touch the end of a buffer, contractions, prefixes and added tokens sliding across the seam."""
    import make_hf_fixture2 as m
    from xllm_service_b200 import workload
    _, tok, hf = setup
    rnd = random.Random(8)
    texts = []
    for n, seed in ((600, 1), (2500, 2)):
        texts.append(" ".join(workload.sentences(1, (n, n), seed=seed)).encode())
    for k in range(80):
        texts.append("".join(rnd.choice(m.ALPHABET) for _ in range(rnd.randrange(800, 5000))).encode())
    texts.append(("1234567" * 900).encode())                   # one 6300-digit run: K-digit tokens across many buffers
    texts.append(("a1" * 4000).encode())
    texts.append(("x \n \n  y!\n\n" * 500).encode())
    texts.append(("  \n" * 30 + "z" + " " * 300 + "\n" + " " * 299 + "q").encode())
    texts.append(("don't WE'LL they'RE x!y " * 400).encode())
    texts.append(("日本語 ١٢٣٤ " * 900).encode())
    texts.append(("x <|endoftext|>\n" * 600).encode())
    for pad in range(1550, 1600):                               # slide interesting contexts across the seam
        texts.append(("ab " * (pad // 3) + "x" * (pad % 3) + " \n \n!'LL 12345\t!a<|endoftext|> é　　y").encode())
    got, status = _encode_all(tok, texts)
    for i, (t, a) in enumerate(zip(texts, got)):
        assert status[i] == 0, (i, len(t), t[:30])
        assert a == hf.prefix_ids + hf.encode(t).tolist() + hf.suffix_ids, (i, len(t), t[:30])


def test_nfc_is_proved_or_refused(setup):
    """normalizer NFC (Qwen2 style): a request passes only when every char is NFC-inert, i.e. NFC provably leaves it
    unchanged; anything else is refused per request (XLLM_ERR_UNSUPPORTED), never tokenised un-normalised."""
    style, tok, hf = setup
    texts = ["plain ascii", "pr\u00e9compos\u00e9 caf\u00e9", "e\u0301 decomposed", "\u212b angstrom sign",
             "\ud55c\uad6d\uc5b4 precomposed", "\u1112\u1161 jamo", "ok " * 1000 + "\u0301", "fine again"]
    got, status = _encode_all(tok, [t.encode() for t in texts])
    if style == "hf_qwen2_style":
        assert status.tolist() == [0, 0, -5, -5, 0, -5, -5, 0]
    else:
        assert (status == 0).all()                               # no normalizer: nothing to prove
    for t, a, st in zip(texts, got, status):
        if st == 0:
            assert a == hf.prefix_ids + hf.encode(t.encode()).tolist() + hf.suffix_ids


def test_ignore_merges_and_template(setup):
    style, tok, hf = setup
    got, status = _encode_all(tok, [b" xyzzyplugh", b"!!!!", b"        ", b"\n \n", b""])
    assert (status == 0).all()
    if style == "hf_llama3_style":
        assert got == [[0, 8000], [0, 8001], [0, 8002], [0, 8003], [0]]   # BOS + the entries no merge produces
    else:
        assert all(len(g) > 1 for g in got[:3]) and got[4] == []
