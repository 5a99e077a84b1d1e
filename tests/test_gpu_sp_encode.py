"""GPU parity of the SentencePiece-BPE encode kernel (csrc/sp_encode.cu) through the C-ABI:
bit-exact token ids against (1) the committed libsentencepiece goldens and (2) the CPU oracle
on seeded random / adversarial inputs."""
import json
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(__file__)
MODEL_DIR = os.path.join(HERE, "golden", "sp_bpe_8k")
GOLD = os.path.join(HERE, "golden", "sp_bpe_8k_goldens.json")


@pytest.fixture(scope="module")
def tok():
    import xllm_service_b200 as x
    h = x.Ingest(tokenizer_path=MODEL_DIR)
    yield h
    h.close()


@pytest.fixture(scope="module")
def sp_oracle(oracle):
    return oracle.SentencePieceOracle(MODEL_DIR)


def _encode_all(tok, texts, stride=None):
    from xllm_service_b200 import workload
    b = workload.pack_prompts(texts)
    if stride is None:
        stride = max(16, 14 * max((len(t) for t in texts), default=0) + 8)  # U+FDFA expands 13x
    ids, n_ids, status = tok.encode_batch(b.text, b.offsets, stride)
    return [ids[i, :n_ids[i]].tolist() for i in range(len(texts))], status


def test_goldens(tok):
    with open(GOLD) as f:
        g = json.load(f)
    texts = [bytes.fromhex(c["text"]) for c in g["cases"]]
    got, status = _encode_all(tok, texts)
    assert (status == 0).all()
    bad = [(texts[i][:40], got[i][:12], g["cases"][i]["ids"][:12]) for i in range(len(texts))
           if got[i] != g["cases"][i]["ids"]]
    assert not bad, bad[:5]


def test_vocab_size(tok):
    assert tok.vocab_size() == 8000


def test_fuzz_vs_oracle(tok, sp_oracle):
    rnd = random.Random(2024)
    alphabet = list("abcdefghijklmnopqrstuvwxyz   \t\n.,!?0123456789") + [
        "é", "日", "ﬁ", "①", "▁", "\U0001F600", "́", " ", "Ａ",
        "ｶﾞ", "　", "​", "ǅ", "㍿", "�", "ﷺ"]
    texts = []
    for i in range(1500):
        k = rnd.randrange(0, 300)
        if i % 3 == 0:
            texts.append(bytes(rnd.getrandbits(8) for _ in range(k)))
        else:
            texts.append("".join(rnd.choice(alphabet) for _ in range(k)).encode())
    got, status = _encode_all(tok, texts)
    assert (status == 0).all()
    for t, g in zip(texts, got):
        assert g == sp_oracle.encode(t).tolist(), t


def test_long_prompts_and_windows(tok, sp_oracle):
    from xllm_service_b200 import workload
    vocab = workload.make_vocabulary()
    texts = [s.encode() for s in workload.sentences(24, (700, 1100), seed=11, vocabulary=vocab)]
    # medium-long words (17..512 chars) take the cooperative path
    rnd = random.Random(3)
    for n in (15, 16, 17, 33, 34, 64, 100, 257, 500):
        texts.append(("".join(rnd.choice("abcdefgh") for _ in range(n)) + " tail " +
                      "".join(rnd.choice("xyz") for _ in range(n))).encode())
    texts.append((" ".join("w%d" % i for i in range(3000))).encode())
    got, status = _encode_all(tok, texts)
    assert (status == 0).all()
    for t, g in zip(texts, got):
        assert g == sp_oracle.encode(t).tolist(), t[:60]


def test_truncation(tok, sp_oracle):
    t = b"hello world this is a test of truncation"
    full = sp_oracle.encode(t).tolist()
    got, status = _encode_all(tok, [t], stride=4)
    assert status[0] == 1  # XLLM_ENC_TRUNCATED
    ids, n_ids, _ = tok.encode_batch(np.frombuffer(t, np.uint8), np.array([0, len(t)], np.int64), 4)
    assert n_ids[0] == len(full) and ids[0].tolist() == full[:4]


def test_very_long_words_take_the_scratch_path(tok, sp_oracle):
    """Whitespace-free runs longer than the shared-memory paths hold (> 512 chars / > 2 KB) are streamed
    through the global scratch slots and must still be bit-exact."""
    rnd = random.Random(17)
    letters = "abcdefghijklmnopqrstuvwxyz"
    texts = [
        b"a" * 5000,
        ("".join(rnd.choice(letters) for _ in range(20000))).encode(),
        ("short words then " + "".join(rnd.choice(letters) for _ in range(3000)) + " and after it more words").encode(),
        ("".join(rnd.choice(letters) for _ in range(1500)) + " " +
         "".join(rnd.choice(letters) for _ in range(2500))).encode(),
        ("日本語のテキスト" * 400).encode(),                      # 3200 unknown chars, byte fallback, no spaces
        ("x" * 700 + "▁" + "y" * 900 + "▁▁" + "z" * 600).encode(),  # literal U+2581 splits long runs
        ("ab" * 1000 + "   ").encode(),                              # long word, then trailing spaces
        ("   " + "q" * 513).encode(),                                # just past the cooperative limit
        (" ".join("w" * n for n in (511, 512, 513, 600, 16, 17, 1)) ).encode(),
        bytes(rnd.choice(b"abcdefgh\xc3\xa9") for _ in range(4000)),
    ]
    got, status = _encode_all(tok, texts)
    assert (status == 0).all(), status
    for t, g in zip(texts, got):
        assert g == sp_oracle.encode(t).tolist(), t[:50]


def test_scratch_capacity_failure_is_loud(sp_oracle):
    import xllm_service_b200 as x
    os.environ["XLLM_SP_LONG_CAP"] = "4096"
    try:
        h = x.Ingest(tokenizer_path=MODEL_DIR)
    finally:
        del os.environ["XLLM_SP_LONG_CAP"]
    big = b"z" * 6000
    ok = b"z" * 3000
    got, status = _encode_all(h, [big, ok, b"fine words"])
    assert status[0] == -6          # XLLM_ERR_CAPACITY: never a silent wrong answer
    assert status[1] == 0 and got[1] == sp_oracle.encode(ok).tolist()
    assert status[2] == 0
    h.close()


def test_empty_batch_and_empty_prompts(tok):
    got, status = _encode_all(tok, [b"", b" ", b"a", b""])
    assert got[0] == [] and got[1] == [] and got[3] == [] and len(got[2]) >= 1
    assert (status == 0).all()
