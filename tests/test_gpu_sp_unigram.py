"""GPU parity of the SentencePiece Unigram backend (csrc/sp_encode.cu unigram_word: upstream's EncodeOptimized
Viterbi, one word at a time from the running float score) against committed pip-sentencepiece goldens
(tests/golden/make_sp_unigram_fixture.py) and the CPU oracle (sentencepiece_tokenizer.cpp:115-168 ->
sp_processor_.Encode; oracle/sp_oracle.cc encode_unigram)."""
import json
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(__file__)
GOLD = os.path.join(HERE, "golden", "sp_unigram_goldens.json")
MODELS = ["sp_unigram_4k", "sp_unigram_4k_bf"]


def _encode_all(tok, texts):
    from xllm_service_b200 import workload
    b = workload.pack_prompts(texts)
    stride = max(16, 3 * max((len(t) for t in texts), default=0) + 8)
    ids, n_ids, status = tok.encode_batch(b.text, b.offsets, stride)
    return [ids[i, :n_ids[i]].tolist() for i in range(len(texts))], status


@pytest.fixture(scope="module", params=MODELS)
def setup(request, oracle):
    import xllm_service_b200 as x
    d = os.path.join(HERE, "golden", request.param)
    h = x.Ingest(tokenizer_path=d)
    yield request.param, h, oracle.SentencePieceOracle(d)
    h.close()


def test_goldens(setup):
    name, tok, _ = setup
    with open(GOLD) as f:
        cases = json.load(f)["cases"][name]
    texts = [bytes.fromhex(c["text"]) for c in cases]
    got, status = _encode_all(tok, texts)
    assert (status == 0).all()
    bad = [(t[:40], a[:12], c["ids"][:12]) for t, a, c in zip(texts, got, cases) if a != c["ids"]]
    assert not bad, (len(bad), bad[:5])


def test_fuzz_and_long_texts_vs_oracle(setup):
    """Prompts of up to 40 KB: the running score reaches -10^5, where float rounding decides close calls — the
    kernel has to carry it exactly as upstream does."""
    from xllm_service_b200 import workload
    _, tok, sp = setup
    rnd = random.Random(13)
    alphabet = list("abcdefghijklmnopqrstuvwxyz   ") + ["é", "日", " ", "\t", "Q", "7", "▁"]
    texts = ["".join(rnd.choice(alphabet) for _ in range(rnd.randrange(0, 80))).encode() for _ in range(1500)]
    texts += [b"\xff\xfe broken \xe6\x97", b"   ", b"a" * 400, ("ab" * 250 + " tail").encode()]
    for n, seed in ((50, 1), (600, 2), (2500, 3), (9000, 4)):
        texts.append(" ".join(workload.sentences(1, (n, n), seed=seed)).encode())
    got, status = _encode_all(tok, texts)
    assert (status == 0).all()
    bad = [(len(t), t[:30]) for t, a in zip(texts, got) if a != sp.encode(t).tolist()]
    assert not bad, (len(bad), bad[:5])


def test_a_word_beyond_the_lattice_is_reported(setup):
    _, tok, sp = setup
    ok = b"z" * 500                      # 3 + 500 bytes with the dummy prefix: fits
    big = b"z" * 600
    got, status = _encode_all(tok, [ok, big, b"fine words"])
    assert status.tolist() == [0, -6, 0]  # XLLM_ERR_CAPACITY: never a silent wrong answer
    assert got[0] == sp.encode(ok).tolist() and got[2] == sp.encode(b"fine words").tolist()


def test_pipeline_with_unigram_tokenizer(setup):
    """The whole ingest path (tokenize -> block hash) with a Unigram model."""
    from xllm_service_b200 import workload
    _, tok, sp = setup
    texts = [s.encode() for s in workload.sentences(300, (5, 200), seed=77)]
    b = workload.pack_prompts(texts)
    out = tok.ingest_batch(b.text, b.offsets, 512, want_match=False)
    assert (out["status"] == 0).all()
    ref_ids, ref_n = sp.encode_batch(b.text, b.offsets, 512)
    assert (out["n_ids"] == ref_n).all()
    from oracle import oracle as o
    for r in range(len(texts)):
        assert (out["ids"][r, :ref_n[r]] == ref_ids[r, :ref_n[r]]).all(), r
        want = o.block_hash_chain(ref_ids[r, :ref_n[r]])
        assert (out["keys"][r, :want.shape[0]] == want).all(), r
