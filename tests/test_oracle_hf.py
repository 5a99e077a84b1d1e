"""Pins the CPU oracle of the reference's HF-tokenizers backend (oracle/hf_bpe_oracle.cc: added-token split,
ByteLevel GPT-2 regex, BPE by merge rank) against upstream `tokenizers`: committed goldens minted with pip
tokenizers 0.22.2 (tests/golden/make_hf_fixture.py) + live fuzz / Unicode sweep when the wheel is importable.
Reference call sites: xllm_service/tokenizer/fast_tokenizer.cpp:20-30, tokenizers/src/lib.rs:83-99."""
import json
import os
import random

import pytest

HERE = os.path.dirname(__file__)
MODEL_DIR = os.path.join(HERE, "golden", "hf_bpe_8k")
GOLD = os.path.join(HERE, "golden", "hf_bpe_goldens.json")


@pytest.fixture(scope="module")
def hf(oracle):
    return oracle.HfBpeOracle(MODEL_DIR)


def test_goldens(hf):
    with open(GOLD) as f:
        g = json.load(f)
    assert g["vocab_size"] == hf.vocab_size == 8000 and len(g["cases"]) > 350
    for c in g["cases"]:
        t = bytes.fromhex(c["text"])
        assert hf.encode(t).tolist() == c["ids"], t[:40]


def test_contract_details(hf):
    assert hf.encode(b"").size == 0
    assert hf.encode(b"<|endoftext|>").tolist() == [0]                    # added token matched on the raw text
    assert hf.encode(b"a<|endoftext|>b").tolist()[1] == 0
    assert hf.encode(b"\xff broken") is None                               # Rust &str cannot hold this: lib.rs:91 panics
    # "\s+(?!\S)": the last space of a run attaches to the following word
    assert hf.encode(b"x   y").tolist() == hf.encode(b"x").tolist() + hf.encode(b"  ").tolist() + hf.encode(b" y").tolist()


def test_live_against_tokenizers_wheel(hf):
    tokenizers = pytest.importorskip("tokenizers")
    tok = tokenizers.Tokenizer.from_file(os.path.join(MODEL_DIR, "tokenizer.json"))
    rnd = random.Random(4)
    alphabet = list("abcdefghij  \t\n'.,!?012") + ["é", "日", "Σ", "١", " ", "　", "\U0001F600", "'s", "'re",
                                                   " '", "<|endoftext|>", "\r\n", " ", "_", "²", "", " "]
    for _ in range(2000):
        s = "".join(rnd.choice(alphabet) for _ in range(rnd.randrange(0, 70)))
        assert hf.encode(s.encode()).tolist() == tok.encode(s).ids, repr(s)
    for cp in list(range(0x80, 0x2100, 3)) + list(range(0x2E80, 0x3100, 5)) + list(range(0x1F600, 0x1F608)):
        s = "a" + chr(cp) + "b " + chr(cp) + "1"
        assert hf.encode(s.encode()).tolist() == tok.encode(s).ids, hex(cp)


# ---------------------------------------------------------------- the newer layouts (Llama-3 / Qwen2 style)
GOLD2 = os.path.join(HERE, "golden", "hf_cl100k_goldens.json")
STYLES = ["hf_llama3_style", "hf_qwen2_style"]


@pytest.mark.parametrize("style", STYLES)
def test_cl100k_family_goldens(oracle, style):
    """Split(cl100k-family regex) + ByteLevel(use_regex=false); ignore_merges + BOS template (Llama-3 style) and
    \\p{N} + NFC (Qwen2 style): goldens minted by tests/golden/make_hf_fixture2.py."""
    h = oracle.HfBpeOracle(os.path.join(HERE, "golden", style))
    assert h.pattern == 2 and h.digits == (3 if style == "hf_llama3_style" else 1)
    assert h.ignore_merges == (style == "hf_llama3_style") and h.nfc == (style == "hf_qwen2_style")
    with open(GOLD2) as f:
        cases = json.load(f)["cases"][style]
    assert len(cases) > 350
    for c in cases:
        t = bytes.fromhex(c["text"])
        assert h.prefix_ids + h.encode(t).tolist() + h.suffix_ids == c["ids"], t[:40]


@pytest.mark.parametrize("style", STYLES)
def test_cl100k_family_live(oracle, style):
    tokenizers = pytest.importorskip("tokenizers")
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_hf_fixture2 as m
    h = oracle.HfBpeOracle(os.path.join(HERE, "golden", style))
    tok = tokenizers.Tokenizer.from_file(os.path.join(HERE, "golden", style, "tokenizer.json"))
    rnd = random.Random(21)
    for _ in range(2500):
        s = "".join(rnd.choice(m.ALPHABET) for _ in range(rnd.randrange(0, 50)))
        assert h.prefix_ids + h.encode(s.encode()).tolist() + h.suffix_ids == tok.encode(s).ids, repr(s)


def test_deepseek_v3_layout_goldens_and_live(oracle):
    """DeepSeek-V3 / R1: Sequence[Split(\\p{N}{1,3}), Split(CJK + kana runs), Split(main regex), ByteLevel], all
    Isolated (tests/golden/make_hf_fixture3.py); goldens from pip `tokenizers`, then live fuzz over an alphabet with
    marks, symbols, control chars, CJK, kana, full-width forms and non-decimal numbers."""
    d = os.path.join(HERE, "golden", "hf_deepseek_style")
    h = oracle.HfBpeOracle(d)
    assert h.pattern == 3 and not h.nfc and not h.ignore_merges and h.prefix_ids == [] and h.suffix_ids == []
    with open(os.path.join(HERE, "golden", "hf_deepseek_goldens.json")) as f:
        cases = json.load(f)["cases"]
    assert len(cases) > 450
    for c in cases:
        t = bytes.fromhex(c["text"])
        assert h.encode(t).tolist() == c["ids"], t[:40]
    tokenizers = pytest.importorskip("tokenizers")
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_hf_fixture3 as m
    tok = tokenizers.Tokenizer.from_file(os.path.join(d, "tokenizer.json"))
    rnd = random.Random(31)
    for _ in range(4000):
        s = "".join(rnd.choice(m.ALPHABET) for _ in range(rnd.randrange(0, 60)))
        assert h.encode(s.encode()).tolist() == tok.encode(s).ids, repr(s)


def test_full_unicode_sweep_against_tokenizers_wheel(oracle):
    """Every code point of planes 0-3 and the assigned ones above (456 K) between letters, after a space, before a
    digit and after a newline: the class tables (scripts/gen_unicode_classes.py, Unicode 15.0) must split exactly as
    the regex engine inside pip `tokenizers` does, for both patterns."""
    tokenizers = pytest.importorskip("tokenizers")
    import unicodedata
    for style in ("hf_bpe_8k", "hf_qwen2_style", "hf_deepseek_style"):
        h = oracle.HfBpeOracle(os.path.join(HERE, "golden", style))
        tok = tokenizers.Tokenizer.from_file(os.path.join(HERE, "golden", style, "tokenizer.json"))
        bad = []
        for cp in range(0x80, 0x110000):
            if 0xD800 <= cp < 0xE000:
                continue
            ch = chr(cp)
            if 0x3FFFF < cp < 0xE0000 and unicodedata.category(ch) == "Cn":
                continue                                    # the empty planes
            if h.nfc and unicodedata.normalize("NFC", ch) != ch:
                continue                                    # the wrapper would normalise it away
            s = "a" + ch + "b " + ch + "1\n" + ch
            if style == "hf_deepseek_style":               # \p{P} / \p{S} / \p{M} contexts of its main regex too
                s += "!" + ch + "x " + ch + ch + "\n.\x01" + ch + "z"
            if h.prefix_ids + h.encode(s.encode()).tolist() + h.suffix_ids != tok.encode(s).ids:
                bad.append(hex(cp))
        assert not bad, (style, len(bad), bad[:20])


def test_nfc_inert_bit_against_the_crates_nfc():
    """The device proves "NFC is the identity" from the per-code-point NFC-inert bit of the class table (bit 2).  Every
    inert code point, alone and in random strings, must be left unchanged by the NFC normalizer of pip `tokenizers`
    (the crate the reference links)."""
    tokenizers = pytest.importorskip("tokenizers")
    import re
    import unicodedata
    nfc = tokenizers.normalizers.NFC()
    with open(os.path.join(HERE, "..", "oracle", "unicode_classes.inc")) as f:
        txt = f.read()
    grab = lambda name: [int(x) for x in re.search(name + r"\[\d+\] = \{(.*?)\};", txt, re.S).group(1).replace("\n", "").split(",") if x.strip()]  # noqa: E731
    s1, s2 = grab("kUniStage1"), grab("kUniStage2")
    inert = [cp for cp in range(0x80, 0x110000)
             if not (0xD800 <= cp < 0xE000) and not (s2[s1[cp >> 8] * 256 + (cp & 255)] & 4)]
    assert len(inert) > 1_000_000
    for i in range(0, len(inert), 4096):
        s = "x".join(chr(cp) for cp in inert[i:i + 4096])
        assert nfc.normalize_str(s) == s, hex(inert[i])
    rnd = random.Random(3)
    assigned = [cp for cp in inert if unicodedata.category(chr(cp)) != "Cn"]
    for _ in range(30000):
        s = "".join(chr(rnd.choice(assigned)) for _ in range(rnd.randrange(1, 8)))
        assert nfc.normalize_str(s) == s, [hex(ord(c)) for c in s]
    # and the usual suspects are NOT inert
    for cp in (0x301, 0x1161, 0x11A8, 0x212B, 0x340, 0xF900, 0x9BE):
        assert s2[s1[cp >> 8] * 256 + (cp & 255)] & 4, hex(cp)
