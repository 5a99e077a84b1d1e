#!/usr/bin/env python3
"""Build the synthetic tiktoken fixture: a byte-level BPE of 1 200 merges trained on the Zipf pseudo-word
corpus (plain python BPE trainer below), written in the reference's vocab format
(xllm_service/tokenizer/tiktoken_tokenizer.cpp:115-153: one `base64(token) rank` per line), and freeze
text -> ids goldens from upstream pip tiktoken 0.12.0 run WITHOUT a regex split (`_encode_single_piece`),
which is the regex-less mode the service uses (tiktoken_tokenizer.cpp:236-241).

Outputs (committed):
  tests/golden/tiktoken_1k/tokenizer.model + tokenizer_config.json ("tokenizer_class": "TikTokenTokenizer")
  tests/golden/tiktoken_goldens.json
"""
import base64
import collections
import json
import os
import random
import sys

import tiktoken

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from xllm_service_b200 import workload  # noqa: E402


def train_bpe(corpus_words, n_merges):
    """corpus_words: Counter of byte strings.  Classic BPE on bytes; returns mergeable_ranks."""
    ranks = {bytes([b]): b for b in range(256)}
    words = {w: [bytes([b]) for b in w] for w in corpus_words}
    for _ in range(n_merges):
        pairs = collections.Counter()
        for w, parts in words.items():
            c = corpus_words[w]
            for a, b in zip(parts, parts[1:]):
                pairs[(a, b)] += c
        if not pairs:
            break
        (a, b), _ = max(pairs.items(), key=lambda kv: (kv[1], kv[0]))
        ranks[a + b] = len(ranks)
        for w, parts in words.items():
            i, out = 0, []
            while i < len(parts):
                if i + 1 < len(parts) and parts[i] == a and parts[i + 1] == b:
                    out.append(a + b)
                    i += 2
                else:
                    out.append(parts[i])
                    i += 1
            words[w] = out
    return ranks


def main():
    out_dir = os.path.join(HERE, "tiktoken_1k")
    os.makedirs(out_dir, exist_ok=True)
    sents = workload.sentences(3000, seed=31)
    cnt = collections.Counter()
    for s in sents:
        for i, w in enumerate(s.split(" ")):
            cnt[((" " if i else "") + w).encode()] += 1
    ranks = train_bpe(cnt, 1200)
    # drop a few single bytes so the "part without an entry is skipped" branch (:222-233) is exercised
    for b in (0x00, 0x7F, 0xF5):
        del ranks[bytes([b])]
    with open(os.path.join(out_dir, "tokenizer.model"), "w") as f:
        for tok, r in sorted(ranks.items(), key=lambda kv: kv[1]):
            f.write(base64.b64encode(tok).decode() + " " + str(r) + "\n")
    with open(os.path.join(out_dir, "tokenizer_config.json"), "w") as f:
        json.dump({"tokenizer_class": "TikTokenTokenizer"}, f)
    enc = tiktoken.Encoding("fixture", pat_str=r"[\s\S]+", mergeable_ranks=ranks, special_tokens={})
    rnd = random.Random(5)
    texts = [b"", b"a", b"hello world", b"  two  spaces ", "日本語 テキスト é".encode(), b"x" * 200,
             b"\xff\xfe\x80 raw bytes \xf5\x00\x7f mixed", b"ab" * 300]
    texts += [s.encode() for s in workload.sentences(60, (3, 60), seed=8)]
    for _ in range(60):
        texts.append(bytes(rnd.choice(b"abcdefghij  \n\xc3\xa9\xe6\x97\xa5\x00\x7f") for _ in range(rnd.randrange(1, 120))))
    gold = []
    missing = {0x00, 0x7F, 0xF5}
    for t in texts:
        if missing & set(t):
            continue  # upstream panics on a byte without a rank; the reference logs and skips it (tested separately)
        ids = enc._encode_single_piece(t) if t else []
        gold.append({"text": t.hex(), "ids": ids})
    with open(os.path.join(HERE, "tiktoken_goldens.json"), "w") as f:
        json.dump({"tiktoken_version": tiktoken.__version__, "n_ranks": len(ranks), "cases": gold}, f,
                  separators=(",", ":"))
    print("ranks", len(ranks), "cases", len(gold))


if __name__ == "__main__":
    main()
