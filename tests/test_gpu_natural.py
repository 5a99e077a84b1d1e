"""The natural-text goldens (tests/golden/natural_goldens.json, see tests/test_oracle_natural.py) through the CUDA
tokenizers: SentencePiece BPE 32 000 pieces and HF byte-level BPE with 128 471 entries — the non-SMALL kernel variants
(ids do not fit 16 bits: 12-byte pair state, 4-id memo payload) at their real vocabulary size, with the memo on and
off, plus a slice of real text far longer than any golden against the CPU oracle."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(__file__)
GOLD = os.path.join(HERE, "golden", "natural_goldens.json")


def _encode_all(h, texts, stride):
    from xllm_service_b200 import workload
    b = workload.pack_prompts(texts)
    ids, n_ids, status = h.encode_batch(b.text, b.offsets, stride)
    assert (status == 0).all(), np.unique(status)
    return [ids[i, :n_ids[i]].tolist() for i in range(len(texts))]


@pytest.mark.parametrize("memo", ["on", "off", "warm"])
@pytest.mark.parametrize("model,key", [("sp_natural_32k", "sp"), ("hf_natural_128k", "hf")])
def test_goldens(model, key, memo):
    import xllm_service_b200 as x
    with open(GOLD) as f:
        cases = json.load(f)["cases"]
    if memo == "off":
        os.environ["XLLM_SP_MEMO_SLOTS"] = "0"
    if memo == "warm":
        os.environ["XLLM_SP_WARM"] = "1"      # the warm-up kernels, which exist for exactly this kind of text
    try:
        h = x.Ingest(tokenizer_path=os.path.join(HERE, "golden", model))
    finally:
        os.environ.pop("XLLM_SP_MEMO_SLOTS", None)
        os.environ.pop("XLLM_SP_WARM", None)
    texts = [bytes.fromhex(c["text"]) for c in cases]
    got = _encode_all(h, texts, 8192)
    bad = [i for i, (g, c) in enumerate(zip(got, cases)) if g != c[key]]
    assert not bad, (len(bad), bad[:5])
    h.close()


@pytest.mark.parametrize("warm", ["plain", "warm"])
@pytest.mark.parametrize("model", ["sp_natural_32k", "hf_natural_128k"])
def test_long_real_text_vs_oracle(oracle, model, warm):
    """~3 MB of this image's own source files cut into 16 KB prompts (what bench.py's natural-text line runs), through
    the plain kernels and through the warm-up kernels."""
    import xllm_service_b200 as x
    from xllm_service_b200 import workload
    corpus = workload.natural_corpus(3 << 20)
    pb = workload.cut_prompts(corpus, 16384)
    d = os.path.join(HERE, "golden", model)
    if warm == "warm":
        os.environ["XLLM_SP_WARM"] = "1"
    try:
        h = x.Ingest(tokenizer_path=d)
    finally:
        os.environ.pop("XLLM_SP_WARM", None)
    ids, n_ids, status = h.encode_batch(pb.text, pb.offsets, 16384)
    assert (status == 0).all()
    if model.startswith("sp"):
        S = oracle.SentencePieceOracle(d)
        enc = lambda t: S.encode(t).tolist()   # noqa: E731
    else:
        H = oracle.HfBpeOracle(d)
        enc = lambda t: H.prefix_ids + H.encode(t).tolist() + H.suffix_ids   # noqa: E731
    bad = [i for i in range(pb.n) if ids[i, :n_ids[i]].tolist() != enc(pb.prompt(i))]
    assert not bad, (len(bad), bad[:5])
    h.close()
