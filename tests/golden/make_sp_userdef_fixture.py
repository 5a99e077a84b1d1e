#!/usr/bin/env python3
"""SentencePiece models WITH user-defined symbols (BPE + byte fallback, and Unigram; vocab 2000) + text -> ids goldens
from upstream libsentencepiece (pip sentencepiece 0.2.1), for the oracle's restatement of the PrefixMatcher path
(normalizer.cc NormalizePrefix, bpe_model.cc frozen symbols, unigram_model.cc user-defined score).  The product
still refuses such models at load; the oracle and its vectors come first.

Outputs (committed): tests/golden/sp_userdef_bpe/tokenizer.model, tests/golden/sp_userdef_unigram/tokenizer.model,
tests/golden/sp_userdef_goldens.json
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

UD = ["<sep>", "<b>", "</b>", "[CLS]", "▁▁▁▁", "<0>", "foo bar", "日本"]


def main():
    rnd = random.Random(23)
    alphabet = list("abcdefghij   ") + UD + ["<", ">", "se", "p>", "<se", "é", "日", "本", "    ", "foo",
                                            " bar", "[CLS", "]"]
    texts = (["", "<sep>", "a<sep>b", "<sep><sep>", "x <b>bold</b> y", "foo bar", "foo  bar", "    x", "日本語",
              "<se p>", "[CLS] hello [CLS]", " <sep> ", "<0><0>"] + workload.sentences(60, (1, 40), seed=3) +
             ["".join(rnd.choice(alphabet) for _ in range(rnd.randrange(0, 40))) for _ in range(400)])
    out = {}
    for name, mt in (("sp_userdef_bpe", "bpe"), ("sp_userdef_unigram", "unigram")):
        model = io.BytesIO()
        spm.SentencePieceTrainer.train(sentence_iterator=iter(workload.sentences(20000, seed=4321)), model_writer=model,
                                       model_type=mt, vocab_size=2000, character_coverage=1.0,
                                       byte_fallback=(mt == "bpe"), user_defined_symbols=UD, minloglevel=2)
        os.makedirs(os.path.join(HERE, name), exist_ok=True)
        with open(os.path.join(HERE, name, "tokenizer.model"), "wb") as f:
            f.write(model.getvalue())
        sp = spm.SentencePieceProcessor(model_proto=model.getvalue())
        out[name] = [{"text": t.encode("utf-8").hex(), "ids": sp.encode(t)} for t in texts]
    with open(os.path.join(HERE, "sp_userdef_goldens.json"), "w") as f:
        json.dump({"sentencepiece_version": spm.__version__, "cases": out}, f, separators=(",", ":"))
    print({k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
