#!/usr/bin/env python3
"""The newer HF `tokenizer.json` layouts on top of the trained vocabulary of hf_bpe_8k (make_hf_fixture.py):

  hf_llama3_style/  Split(cl100k-family regex with \\p{N}{1,3}, Isolated) + ByteLevel(use_regex=false),
                    ignore_merges = true (extra vocabulary entries no merge produces), BOS template
  hf_qwen2_style/   the same regex with a single \\p{N}, normalizer NFC, no template

and freezes text -> ids goldens from upstream HF `tokenizers` (pip 0.22.2; the reference links crate 0.21,
xllm_service/tokenizer/tokenizers/Cargo.toml:11) through the call the service makes
(fast_tokenizer.cpp:20-30 -> tokenizers_encode(text, add_special_tokens = 1)).

Outputs (committed): tests/golden/hf_llama3_style/tokenizer.json, tests/golden/hf_qwen2_style/tokenizer.json,
tests/golden/hf_cl100k_goldens.json
"""
import json
import os
import random
import sys

from tokenizers import Tokenizer

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from xllm_service_b200 import workload  # noqa: E402

P3 = (r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+"
      r"|\s+(?!\S)|\s+")
P1 = P3.replace(r"\p{N}{1,3}", r"\p{N}")
G = "Ġ"   # byte-level spelling of 0x20
C = "Ċ"   # byte-level spelling of 0x0A
EXTRA = [G + "xyzzyplugh", "!!!!", G * 8, C + G + C]   # in the vocabulary, produced by no merge


def build(base, name, pat, ignore_merges, nfc, template):
    d = json.loads(json.dumps(base))
    d["pre_tokenizer"] = {"type": "Sequence", "pretokenizers": [
        {"type": "Split", "pattern": {"Regex": pat}, "behavior": "Isolated", "invert": False},
        {"type": "ByteLevel", "add_prefix_space": False, "trim_offsets": True, "use_regex": False}]}
    d["model"]["ignore_merges"] = ignore_merges
    if ignore_merges:
        nxt = max(d["model"]["vocab"].values()) + 1
        for t in EXTRA:
            d["model"]["vocab"][t] = nxt
            nxt += 1
    d["normalizer"] = {"type": "NFC"} if nfc else None
    bl = {"type": "ByteLevel", "add_prefix_space": True, "trim_offsets": False, "use_regex": True}
    if template:
        d["post_processor"] = {"type": "Sequence", "processors": [bl, {
            "type": "TemplateProcessing",
            "single": [{"SpecialToken": {"id": "<|endoftext|>", "type_id": 0}}, {"Sequence": {"id": "A", "type_id": 0}}],
            "pair": [{"SpecialToken": {"id": "<|endoftext|>", "type_id": 0}}, {"Sequence": {"id": "A", "type_id": 0}},
                     {"SpecialToken": {"id": "<|endoftext|>", "type_id": 1}}, {"Sequence": {"id": "B", "type_id": 1}}],
            "special_tokens": {"<|endoftext|>": {"id": "<|endoftext|>", "ids": [0], "tokens": ["<|endoftext|>"]}}}]}
    else:
        d["post_processor"] = bl
    os.makedirs(os.path.join(HERE, name), exist_ok=True)
    path = os.path.join(HERE, name, "tokenizer.json")
    with open(path, "w", encoding="utf-8") as f:
        json.dump(d, f, ensure_ascii=False, separators=(",", ":"))
    return path


ALPHABET = list("abcdefghij  \t\n'.,!?012") + [
    "é", "日", "Σ", "١", " ", "　", "\U0001F600", "'s", "'RE", " '", "<|endoftext|>", "\r\n",
    " ", "_", "²", "", "'ſ", "'Ll", "'D", "\n\n", " \n", "\n ", "!\n", "12345", " 7", "x!y", " !",
    "\t!", "\r", "  ", " xyzzyplugh", "!!!!", "        ", "\n \n", "1", "٣٤٥٦"]


def texts(rnd):
    t = ["", " ", "  ", "a", "Hello world", "Hello  world", "Hello world's  test\n\n 123 don't   x",
         "I'm you're we've they'll he'd it's can't 'tis 'Twas O'Neil DON'T I'M WE'LL", "'S'T'RE'VE'M'LL'D", "''''",
         "x'ſa 'Sx 'RE 'Re 'lL", "!!a x!a x !a  !a\t!a", "12345 1234567 1 12 123 1234", "a \n \n b", "!\n\n x",
         "a   \n  b", "  x", "\t\tx", "a  b", "x\n\ny", "'abc", " 's", "a\r\n\r\nb",
         "१२३४ a１２３４", "a \n", "a\n ", "a  ", "\n\n\nabc", "!!\n\n\n\n",
         "a  \n\n  \n  b", "x <|endoftext|> y", "a<|endoftext|>\nb", "tabs\tand\nnewlines\r\n\r\n  end  ",
         "numbers 3.14 1,000 1e10 x2y 2024-01-01", "(a) [b] {c} <d> a+b=c #tag @user $5",
         " xyzzyplugh xyzzyplugh!!!!", "        x", "\n \n", "mixed123abc456 αβγ123 日本語123",
         "café naïve Ünïcödé", "x" * 300, "ab " * 200, "1234567890" * 30, "!?" * 100,
         " " * 40 + "z", "\n" * 20, " \n" * 20, "\n " * 20]
    out = list(t)
    for _ in range(300):
        out.append("".join(rnd.choice(ALPHABET) for _ in range(rnd.randrange(1, 80))))
    out += workload.sentences(60, (3, 80), seed=43)
    return out


def main():
    with open(os.path.join(HERE, "hf_bpe_8k", "tokenizer.json"), encoding="utf-8") as f:
        base = json.load(f)
    rnd = random.Random(3)
    cases = texts(rnd)
    gold = {}
    for name, pat, im, nfc, tmpl in (("hf_llama3_style", P3, True, False, True),
                                     ("hf_qwen2_style", P1, False, True, False)):
        tok = Tokenizer.from_file(build(base, name, pat, im, nfc, tmpl))
        gold[name] = [{"text": s.encode("utf-8").hex(), "ids": tok.encode(s, add_special_tokens=True).ids}
                      for s in cases]
    import tokenizers
    with open(os.path.join(HERE, "hf_cl100k_goldens.json"), "w") as f:
        json.dump({"tokenizers_version": tokenizers.__version__, "cases": gold}, f, separators=(",", ":"))
    print({k: len(v) for k, v in gold.items()})


if __name__ == "__main__":
    main()
