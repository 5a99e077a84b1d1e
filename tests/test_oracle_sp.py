"""Pins the CPU oracle's SentencePiece-BPE encoder (oracle/sp_oracle.cc) against upstream libsentencepiece:
(1) the committed goldens minted with pip sentencepiece 0.2.1 (tests/golden/make_sp_fixture.py) —
    text -> ids and text -> normalized string, incl. adversarial UTF-8;
(2) live against the sentencepiece wheel when it is importable on this machine.
Reference call sites: xllm_service/tokenizer/sentencepiece_tokenizer.cpp:47-50,115-168."""
import json
import os
import random

import numpy as np
import pytest

HERE = os.path.dirname(__file__)
MODEL_DIR = os.path.join(HERE, "golden", "sp_bpe_8k")
GOLD = os.path.join(HERE, "golden", "sp_bpe_8k_goldens.json")


@pytest.fixture(scope="module")
def sp(oracle):
    return oracle.SentencePieceOracle(MODEL_DIR)


def test_goldens_ids_and_normalization(sp):
    with open(GOLD) as f:
        g = json.load(f)
    assert g["vocab_size"] == sp.vocab_size() == 8000
    assert len(g["cases"]) > 400
    for c in g["cases"]:
        t = bytes.fromhex(c["text"])
        assert sp.normalize(t).hex() == c["norm"], t[:40]
        assert sp.encode(t).tolist() == c["ids"], t[:40]


def test_contract_details(sp):
    assert sp.encode(b"").size == 0                      # sentencepiece_tokenizer.cpp:117-120
    assert sp.encode(b"   ").size == 0                   # all-whitespace normalises to nothing
    a = sp.encode(b"hello world").tolist()
    assert sp.encode(b"  hello   world  ").tolist() == a  # remove_extra_whitespaces + dummy prefix
    # byte fallback: an unknown char becomes its UTF-8 bytes' <0xNN> pieces (ids 3..258)
    ids = sp.encode("日".encode()).tolist()
    assert ids[1:] == [3 + 0xE6, 3 + 0x97, 3 + 0xA5]
    # invalid UTF-8 -> U+FFFD -> bytes EF BF BD
    assert sp.encode(b"\xff").tolist()[1:] == [3 + 0xEF, 3 + 0xBF, 3 + 0xBD]


def test_batch_threads_equal_single(sp):
    from xllm_service_b200 import workload
    texts = [s.encode() for s in workload.sentences(40, (5, 80), seed=9)] + [b"", b"\xf0\x9f\x98\x80 x"]
    b = workload.pack_prompts(texts)
    ids, n = sp.encode_batch(b.text, b.offsets, 400, n_threads=4)
    for i, t in enumerate(texts):
        assert ids[i, :n[i]].tolist() == sp.encode(t).tolist()


def test_rejects_non_bpe_models(oracle, tmp_path):
    with pytest.raises(ValueError):
        oracle.SentencePieceOracle(str(tmp_path / "missing.model"))
    bad = tmp_path / "tokenizer.model"
    bad.write_bytes(b"\x00\x01garbage")
    with pytest.raises(ValueError):
        oracle.SentencePieceOracle(str(bad))


def test_live_against_sentencepiece_wheel(sp):
    spm = pytest.importorskip("sentencepiece")
    ref = spm.SentencePieceProcessor(model_file=os.path.join(MODEL_DIR, "tokenizer.model"))
    rnd = random.Random(77)
    alphabet = list("abcdefghijklmnopqrstuvwxyz   \t\n.,!?0123456789") + [
        "é", "日", "ﬁ", "①", "▁", "\U0001F600", "́", " ", "Ａ", "ﷺ", "�"]
    for i in range(1500):
        k = rnd.randrange(0, 160)
        t = bytes(rnd.getrandbits(8) for _ in range(k)) if i % 3 == 0 else \
            "".join(rnd.choice(alphabet) for _ in range(k)).encode()
        assert sp.encode(t).tolist() == ref.EncodeAsIds(t), t
        assert sp.normalize(t) == ref.Normalize(t), t


# ---------------------------------------------------------------- Unigram models (oracle first; no device kernel yet)
@pytest.mark.parametrize("name", ["sp_unigram_4k", "sp_unigram_4k_bf"])
def test_unigram_goldens_and_live(oracle, name):
    """unigram_model.cc EncodeOptimized restated in oracle/sp_oracle.cc: committed goldens from pip sentencepiece
    0.2.1 (tests/golden/make_sp_unigram_fixture.py) + live fuzz when the wheel is importable."""
    import json
    import random
    d = os.path.join(HERE, "golden", name)
    S = oracle.SentencePieceOracle(d)
    with open(os.path.join(HERE, "golden", "sp_unigram_goldens.json")) as f:
        cases = json.load(f)["cases"][name]
    assert len(cases) > 400
    for c in cases:
        t = bytes.fromhex(c["text"])
        assert S.encode(t).tolist() == c["ids"], t[:40]
    spm = pytest.importorskip("sentencepiece")
    sp = spm.SentencePieceProcessor(model_file=os.path.join(d, "tokenizer.model"))
    rnd = random.Random(31)
    alphabet = list("abcdefghijklmnopqrstuvwxyz   ") + ["é", "日", " ", "\t", "Q", "7", "▁"]
    for _ in range(1500):
        t = "".join(rnd.choice(alphabet) for _ in range(rnd.randrange(0, 50)))
        assert S.encode(t.encode()).tolist() == sp.encode(t), repr(t)


@pytest.mark.parametrize("name", ["sp_userdef_bpe", "sp_userdef_unigram"])
def test_user_defined_symbols_goldens(oracle, name):
    """USER_DEFINED pieces (PrefixMatcher: matched on the raw text, passed through the normalizer verbatim, frozen in
    BPE, always preferred in Unigram): goldens from pip sentencepiece 0.2.1 (make_sp_userdef_fixture.py).  Oracle
    only — the product refuses these models at load."""
    import json
    import xllm_service_b200 as x
    from xllm_service_b200 import _lib
    d = os.path.join(HERE, "golden", name)
    S = oracle.SentencePieceOracle(d)
    with open(os.path.join(HERE, "golden", "sp_userdef_goldens.json")) as f:
        cases = json.load(f)["cases"][name]
    assert len(cases) > 400
    for c in cases:
        t = bytes.fromhex(c["text"])
        assert S.encode(t).tolist() == c["ids"], t[:40]
    with pytest.raises(x.IngestError) as e:
        _lib.tokenizer_probe(d)
    assert e.value.code == -5


def test_full_unicode_sweep_against_the_wheel(oracle):
    """Every code point (1.1 M) inside a word and after a space through normalisation (the model's nmt_nfkc charsmap),
    BPE and byte fallback: the oracle must agree with upstream libsentencepiece on all of them."""
    spm = pytest.importorskip("sentencepiece")
    d = os.path.join(HERE, "golden", "sp_bpe_8k")
    sp = spm.SentencePieceProcessor(model_file=os.path.join(d, "tokenizer.model"))
    S = oracle.SentencePieceOracle(d)
    bad = []
    for cp in range(1, 0x110000):
        if 0xD800 <= cp < 0xE000:
            continue
        s = "a" + chr(cp) + "b " + chr(cp)
        if S.encode(s.encode()).tolist() != sp.encode(s):
            bad.append(hex(cp))
    assert not bad, (len(bad), bad[:20])
