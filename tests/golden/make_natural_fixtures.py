#!/usr/bin/env python3
"""Tokenizers trained on REAL text for the honest-text bench numbers: the synthetic headline
workload is 20 K lowercase pseudo-words; these two models see source code and prose with punctuation, digits,
newlines, indentation, long identifiers and UTF-8.

  sp_natural_32k/tokenizer.model   SentencePiece BPE, 32 000 pieces, byte fallback, default nmt_nfkc normaliser
  hf_natural_128k/tokenizer.json   HF byte-level BPE (GPT-2 layout), 131 072-entry vocabulary (>= 100 K merges):
                                   ids do not fit 16 bits -> the non-SMALL kernel variants (12-byte pair state,
                                   4-id memo payload)

Corpus: workload.natural_corpus() — the *.py / *.md / *.rst / *.txt / *.h / *.hpp files of this image's site-packages
in sorted order (the first 48 MB for training).  Goldens: a few hundred slices of the same kind of text frozen from
pip sentencepiece / pip tokenizers (tests/golden/natural_goldens.json).
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from xllm_service_b200 import workload  # noqa: E402


def main():
    import sentencepiece as spm
    from tokenizers import Tokenizer, decoders, models, pre_tokenizers, trainers
    corpus = workload.natural_corpus(48 << 20)
    lines = [ln for ln in corpus.decode("utf-8").split("\n") if ln.strip()]
    print("corpus", len(corpus), "bytes,", len(lines), "lines")
    tmp = "/tmp/natural_train.txt"
    with open(tmp, "w", encoding="utf-8") as f:
        f.write("\n".join(lines))
    # ---- SentencePiece BPE 32k
    d = os.path.join(HERE, "sp_natural_32k")
    os.makedirs(d, exist_ok=True)
    spm.SentencePieceTrainer.train(input=tmp, model_prefix=os.path.join(d, "tokenizer"), vocab_size=32000,
                                   model_type="bpe", byte_fallback=True, character_coverage=0.9995,
                                   input_sentence_size=400000, shuffle_input_sentence=True, num_threads=8,
                                   max_sentence_length=16384, train_extremely_large_corpus=False)
    os.remove(os.path.join(d, "tokenizer.vocab"))
    # ---- HF byte-level BPE, 131072 entries
    d2 = os.path.join(HERE, "hf_natural_128k")
    os.makedirs(d2, exist_ok=True)
    tok = Tokenizer(models.BPE())
    tok.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=True)
    tok.decoder = decoders.ByteLevel()
    trainer = trainers.BpeTrainer(vocab_size=131072, special_tokens=["<|endoftext|>"], show_progress=False,
                                  initial_alphabet=pre_tokenizers.ByteLevel.alphabet())
    tok.train_from_iterator(lines, trainer=trainer)
    tok.save(os.path.join(d2, "tokenizer.json"))
    print("hf vocab", tok.get_vocab_size())
    # ---- goldens on held-out text (bytes 48 MB .. 50 MB of the same corpus order)
    held = workload.natural_corpus(50 << 20)[48 << 20:]
    pb = workload.cut_prompts(held, 6000)
    sp = spm.SentencePieceProcessor(model_file=os.path.join(d, "tokenizer.model"))
    cases = []
    for i in range(min(pb.n, 300)):
        t = pb.prompt(i)
        s = t.decode("utf-8", errors="ignore")
        cases.append({"text": s.encode("utf-8").hex(), "sp": sp.encode(s), "hf": tok.encode(s).ids})
    import sentencepiece
    import tokenizers
    with open(os.path.join(HERE, "natural_goldens.json"), "w") as f:
        json.dump({"sentencepiece_version": sentencepiece.__version__, "tokenizers_version": tokenizers.__version__,
                   "cases": cases}, f, separators=(",", ":"))
    print("goldens", len(cases))


if __name__ == "__main__":
    main()
