"""GPU parity of the express tokenizer kernel (csrc/sp_encode.cu: sp_express_kernel / express_run) at its seams,
bit-exact against the CPU oracle through the C-ABI: window boundaries (256-byte windows, 8-byte alignment, byte 255
never consumed), more than 32 words in a window, words of exactly 15 / 16 / 17 bytes (lane path vs cooperative path),
space-like bytes, runs of spaces, truncating rows, hand-over to the buffer-path kernel at every kind of byte the
express rules do not cover (with a pending U+2581, without, right after a word, inside a run of spaces), and the same
prompts with the express kernel switched off (XLLM_SP_EXPRESS=0) as a second witness."""
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(__file__)
MODEL_8K = os.path.join(HERE, "golden", "sp_bpe_8k")
MODEL_32K = os.path.join(HERE, "golden", "sp_natural_32k")


def _encode_all(tok, texts, stride=None):
    from xllm_service_b200 import workload
    b = workload.pack_prompts(texts)
    if stride is None:
        stride = max(16, 14 * max((len(t) for t in texts), default=0) + 8)
    ids, n_ids, status = tok.encode_batch(b.text, b.offsets, stride)
    return [ids[i, :n_ids[i]].tolist() for i in range(len(texts))], status, n_ids


def _check(tok, orc, texts):
    got, status, _ = _encode_all(tok, texts)
    assert (status == 0).all(), status
    bad = [(i, texts[i][:48]) for i in range(len(texts)) if got[i] != orc.encode(texts[i]).tolist()]
    assert not bad, bad[:5]


@pytest.fixture(scope="module", params=[MODEL_8K, MODEL_32K], ids=["bpe8k", "natural32k"])
def pair(request, oracle):
    import xllm_service_b200 as x
    h = x.Ingest(tokenizer_path=request.param)
    yield h, oracle.SentencePieceOracle(request.param)
    h.close()


def _words(rnd, n, lo=1, hi=9, alphabet="abcdefghijklmnopqrstuvwxyz"):
    return ["".join(rnd.choice(alphabet) for _ in range(rnd.randint(lo, hi))) for _ in range(n)]


def test_window_boundaries_and_alignment(pair):
    """Every text length around one and two windows, at every start alignment (the batch packs prompts back to back,
    so a prompt of length L shifts the next one's address by L)."""
    tok, orc = pair
    rnd = random.Random(5)
    base = " ".join(_words(rnd, 400)).encode()
    texts = []
    for n in list(range(0, 40)) + list(range(240, 280)) + list(range(500, 530)) + [767, 768, 769, 1023, 1024, 1025]:
        off = rnd.randrange(0, 64)
        texts.append(base[off:off + n])
    # a word that ends exactly at byte 253 / 254 / 255 / 256 of its window, and a space at byte 255
    for pad in range(244, 262):
        texts.append(b"a" * 3 + b" " + b"b" * (pad - 4) + b" " + b"tail word here")
        texts.append(b"x " * (pad // 2) + b"end")
    _check(tok, orc, texts)


def test_more_than_32_words_in_a_window(pair):
    tok, orc = pair
    texts = [b"a " * 200, b"a b " * 100 + b"cd", b" ".join(b"w" for _ in range(129)), b"ab " * 300,
             b"a  b   c    d " * 40, (b"i " * 31 + b"longerword ") * 12]
    _check(tok, orc, texts)


def test_word_lengths_around_the_lane_limit(pair):
    tok, orc = pair
    rnd = random.Random(9)
    texts = []
    for n in (1, 2, 7, 8, 9, 14, 15, 16, 17, 18, 31, 32, 33, 63, 64, 65, 127, 128, 200, 255, 256, 257, 300, 511, 512):
        w = "".join(rnd.choice("abcdefghij") for _ in range(n))
        texts.append(("head " + w + " tail").encode())
        texts.append((w + " " + w).encode())                 # the same word twice: miss then memo hit (when it fits a key)
        texts.append(("  " + w).encode())
        texts.append((w + "   ").encode())
    texts.append((" ".join("".join(rnd.choice("abc") for _ in range(rnd.choice((3, 15, 16, 40)))) for _ in range(400))).encode())
    _check(tok, orc, texts)


def test_space_like_bytes_and_space_runs(pair):
    tok, orc = pair
    rnd = random.Random(21)
    seps = [" ", "  ", "\t", "\n", "\r\n", " \n ", "\n\n\n", "   \t  "]
    texts = []
    for _ in range(60):
        ws = _words(rnd, rnd.randint(1, 120), 1, 12, "abcdefghijklmnopqrstuvwxyzABC.,;:(){}[]=+-*/_'\"0123456789")
        texts.append("".join(w + rnd.choice(seps) for w in ws).encode())
        texts.append((rnd.choice(seps) + rnd.choice(seps)).encode() + texts[-1])
    texts += [b" ", b"  ", b"\n", b" \n\t ", b" " * 255, b" " * 256, b" " * 257, b" " * 600 + b"x", b"x" + b" " * 600,
              b"\x7f word", b"a\x00b", b"tab\there", b"bell\x07x y z"]
    _check(tok, orc, texts)


def test_hand_over_to_the_buffer_path(pair):
    """A byte the express rules do not cover sends the rest of the request to the buffer-path kernel with the state a
    drain would have left: at every position class, near and far from window boundaries."""
    tok, orc = pair
    rnd = random.Random(33)
    specials = ["é", "日本", "ﬁ", "▁", "　", "​", "\U0001F600", "é", "\xa0", "Ａ"]
    raw = [b"\xff", b"\xc3", b"\xe2\x96", b"\x80abc"]
    texts = []
    for sp in specials:
        for lead in (0, 1, 7, 100, 250, 251, 255, 256, 257, 700):
            prefix = " ".join(_words(rnd, 200))[:lead]
            for glue in ("", " ", "  ", "x", "x "):
                texts.append((prefix + glue + sp + glue + "after it more words follow here").encode())
    for r in raw:
        for lead in (0, 5, 254, 255, 256, 300):
            texts.append((" ".join(_words(rnd, 100))[:lead]).encode() + r + b" and words after")
    # mostly ASCII prose with a few non-ASCII chars late in a long prompt
    body = " ".join(_words(rnd, 3000, 1, 20))
    texts.append((body[:9000] + " naïve café " + body[9000:]).encode())
    _check(tok, orc, texts)


def test_truncating_row_in_the_express_kernel(pair):
    tok, orc = pair
    rnd = random.Random(4)
    t = " ".join(_words(rnd, 500)).encode()
    full = orc.encode(t).tolist()
    for cap in (1, 7, 31, 32, 33, 100, len(full) - 1, len(full), len(full) + 5):
        ids, n_ids, status = tok.encode_batch(np.frombuffer(t, np.uint8), np.array([0, len(t)], np.int64), cap)
        assert n_ids[0] == len(full)
        assert status[0] == (1 if cap < len(full) else 0)
        assert ids[0, :min(cap, len(full))].tolist() == full[:cap]


def test_express_off_gives_the_same_ids(oracle):
    """XLLM_SP_EXPRESS=0 routes everything through the buffer-path kernel: both must agree with the oracle."""
    import xllm_service_b200 as x
    rnd = random.Random(77)
    texts = [(" ".join(_words(rnd, rnd.randint(1, 800), 1, 18))).encode() for _ in range(40)]
    texts += [("mixed é text " * 50).encode(), b"", b" ", b"one"]
    orc = oracle.SentencePieceOracle(MODEL_8K)
    want = [orc.encode(t).tolist() for t in texts]
    os.environ["XLLM_SP_EXPRESS"] = "0"
    try:
        h0 = x.Ingest(tokenizer_path=MODEL_8K)
    finally:
        del os.environ["XLLM_SP_EXPRESS"]
    h1 = x.Ingest(tokenizer_path=MODEL_8K)
    try:
        g0, s0, _ = _encode_all(h0, texts)
        g1, s1, _ = _encode_all(h1, texts)
        assert (s0 == 0).all() and (s1 == 0).all()
        assert g0 == want and g1 == want
    finally:
        h0.close()
        h1.close()
