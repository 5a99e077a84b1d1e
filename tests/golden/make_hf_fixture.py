#!/usr/bin/env python3
"""Train the synthetic HF byte-level BPE fixture (GPT-2 style `tokenizer.json`) and freeze text -> ids
goldens from upstream HF `tokenizers` (pip 0.22.2; the reference pins crate 0.21.0 without a lock file,
xllm_service/tokenizer/tokenizers/Cargo.toml:11 — BPE + ByteLevel behaviour is unchanged between them).

This is the backend the service uses whenever `<tokenizer_path>/tokenizer.json` exists
(xllm_service/tokenizer/tokenizer_factory.cpp:14-19 -> FastTokenizer -> tokenizers_encode with
add_special_tokens = 1, fast_tokenizer.cpp:20-30).

Outputs (committed):
  tests/golden/hf_bpe_8k/tokenizer.json
  tests/golden/hf_bpe_goldens.json
"""
import json
import os
import random
import sys

from tokenizers import Tokenizer, decoders, models, pre_tokenizers, trainers

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from xllm_service_b200 import workload  # noqa: E402


def texts(rnd):
    t = [
        "", " ", "  ", "a", "Hello world", "Hello  world", "Hello world's  test\n\n 123 don't   x",
        "I'm you're we've they'll he'd it's can't 'tis 'Twas O'Neil DON'T I'M", "'s't're've'm'll'd", "''''", "a'b'c",
        "tabs\tand\nnewlines\r\n\r\n  end  ", "trailing spaces   ", "   leading", "x \n y", "x\n y", "x \ny",
        "numbers 1 22 333 4444 3.14 1,000 1e10 x2y", "punct !!! ... --- (a) [b] {c} <d> a+b=c #tag @user $5 50%",
        "mixed123abc456 αβγ123 日本語123", "café naïve Ünïcödé Σίσυφος Привет мир", "日本語のテキスト 中文 한국어",
        "emoji \U0001F600\U0001F44D\U0001F3FD ok", "nbsp here", "em space wide　space", " line sep", "nel",
        "²³¼½ ① Ⅷ ٣ ๓", "a_b __init__ snake_case camelCase", "<|endoftext|>", "a<|endoftext|>b", "x <|endoftext|> y",
        "<|endoftext|><|endoftext|>", "<|endoftext", "<|endoftext|", "|endoftext|>", "text<|endoftext|>" * 3,
        "x" * 300, "ab " * 200, "1234567890" * 30, "!?" * 100, " " * 40 + "z", "\n" * 20, " \n" * 20, "\n " * 20,
    ]
    alphabet = list("abcdefghij  \t\n'.,!?012") + ["é", "日", "Σ", "١", " ", "　", "\U0001F600", "'s", "'re", " '",
                                                   "<|endoftext|>", "\r\n", " ", "_", "²"]
    out = [s for s in t]
    for _ in range(250):
        out.append("".join(rnd.choice(alphabet) for _ in range(rnd.randrange(1, 80))))
    out += workload.sentences(100, (3, 80), seed=41)
    out.append(" ".join(workload.sentences(30, (30, 60), seed=6)))
    return out


def main():
    out_dir = os.path.join(HERE, "hf_bpe_8k")
    os.makedirs(out_dir, exist_ok=True)
    tok = Tokenizer(models.BPE())
    tok.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=True)
    tok.decoder = decoders.ByteLevel()
    trainer = trainers.BpeTrainer(vocab_size=8000, initial_alphabet=pre_tokenizers.ByteLevel.alphabet(),
                                  special_tokens=["<|endoftext|>"], show_progress=False)
    tok.train_from_iterator(workload.sentences(20000, seed=4321), trainer)
    tok.save(os.path.join(out_dir, "tokenizer.json"))
    with open(os.path.join(out_dir, "tokenizer_config.json"), "w") as f:
        json.dump({"tokenizer_class": "GPT2TokenizerFast"}, f)
    rnd = random.Random(2)
    gold = []
    for s in texts(rnd):
        ids = tok.encode(s, add_special_tokens=True).ids  # fast_tokenizer.cpp:24
        gold.append({"text": s.encode("utf-8").hex(), "ids": ids})
    import tokenizers
    with open(os.path.join(HERE, "hf_bpe_goldens.json"), "w") as f:
        json.dump({"tokenizers_version": tokenizers.__version__, "vocab_size": tok.get_vocab_size(), "cases": gold}, f,
                  separators=(",", ":"))
    print("cases", len(gold), "vocab", tok.get_vocab_size())


if __name__ == "__main__":
    main()
