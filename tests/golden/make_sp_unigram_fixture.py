#!/usr/bin/env python3
"""Unigram SentencePiece fixtures (vocab 4000, with and without byte_fallback) + text -> ids goldens from upstream
libsentencepiece (pip sentencepiece 0.2.1), for the oracle's restatement of unigram_model.cc EncodeOptimized
(oracle/sp_oracle.cc encode_unigram).  The device kernel for Unigram is not built yet; the oracle and
its vectors come first.

Outputs (committed): tests/golden/sp_unigram_4k/tokenizer.model, tests/golden/sp_unigram_4k_bf/tokenizer.model,
tests/golden/sp_unigram_goldens.json
"""
import io
import json
import os
import random
import sys

import sentencepiece as spm

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from xllm_service_b200 import workload  # noqa: E402


def main():
    out = {}
    rnd = random.Random(17)
    alphabet = list("abcdefghijklmnopqrstuvwxyz   ") + ["é", "日本", " ", "\t", "xyzzy", "Q", "1",
                                                         "▁", "\U0001F642", "  ", "\n"]
    texts = (["", " ", "a", "hello world", "  leading and trailing  ", "unknown 日本語 chars \U0001F642",
              "x" * 200, "ab " * 100] + workload.sentences(150, (1, 60), seed=9) +
             ["".join(rnd.choice(alphabet) for _ in range(rnd.randrange(0, 60))) for _ in range(350)])
    for name, bf in (("sp_unigram_4k", False), ("sp_unigram_4k_bf", True)):
        model = io.BytesIO()
        spm.SentencePieceTrainer.train(sentence_iterator=iter(workload.sentences(20000, seed=4321)), model_writer=model,
                                       model_type="unigram", vocab_size=4000, character_coverage=1.0, byte_fallback=bf,
                                       minloglevel=2)
        os.makedirs(os.path.join(HERE, name), exist_ok=True)
        with open(os.path.join(HERE, name, "tokenizer.model"), "wb") as f:
            f.write(model.getvalue())
        sp = spm.SentencePieceProcessor(model_proto=model.getvalue())
        out[name] = [{"text": t.encode("utf-8").hex(), "ids": sp.encode(t)} for t in texts]
    with open(os.path.join(HERE, "sp_unigram_goldens.json"), "w") as f:
        json.dump({"sentencepiece_version": spm.__version__, "cases": out}, f, separators=(",", ":"))
    print({k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
