#!/usr/bin/env python3
"""Train the synthetic SentencePiece-BPE fixture model the tests and bench.py use
(vocab 8000, byte_fallback, character_coverage 1.0, trained on 20 000
Zipf-0.9 pseudo-word sentences) and freeze (text -> ids, text -> normalized) golden pairs
from upstream libsentencepiece (pip sentencepiece 0.2.1 wraps it).  The reference holds no
tokenizer fixture of its own.

Outputs (committed):
  tests/golden/sp_bpe_8k/tokenizer.model          the model (read by the oracle and the product)
  tests/golden/sp_bpe_8k/tokenizer_config.json    makes the directory a valid --tokenizer_path
  tests/golden/sp_bpe_8k_goldens.json             text(hex) -> ids / normalized(hex)
Run:  python tests/golden/make_sp_fixture.py
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


def adversarial_texts(rnd):
    t = [
        "", " ", "  ", "a", " a", "a ", "  a  b  ", "hello world", "hello  world", "Hello, World!",
        "\t\ttabs\tand\nnewlines\r\n", "ﬁne ﬂuﬀy ①②③ ｆｕｌｌ",
        "日本語のテキスト", "é é ê ë é",
        "▁leading meta", "trailing meta▁", "mid▁dle ▁▁ x", " nbsp 　ideographic space",
        "emoji \U0001F600\U0001F44D\U0001F3FD zwj \U0001F468‍\U0001F469‍\U0001F467",
        "​zero​width﻿", "à́̂ combining",
        "İstanbul ǅ ß ẞ Σς", "١٢٣ ०१२ ㊙ ㌀ ㍿",
        "x" * 300, "ab " * 200, " " * 50 + "z" + " " * 50,
        "<s> </s> <unk> <0x41> <0xFF>", "\x00\x01\x02 control \x7f chars", "½ ¼ ² ³ ™ © ®",
        "ＡＢＣ ａｂｃ １２３",
        "ｶﾞｷﾞｸﾞ half-width kana", "한국어 조합 한",
        "� replacement char literal", "Ǆ ǆ ﬃ ﬄ ﬅ ﬆ", "a­b soft hyphen",
        " line sep para sep", "nel", " en em thin narrow math",
    ]
    raw = [
        b"\xff\xfe invalid bytes \x80\x81", b"trunc \xe6\x97", b"\xc0\xaf overlong", b"\xed\xa0\x80 surrogate",
        b"\xf4\x90\x80\x80 too big", b"\xe2\x96", b"ok\xe2\x96\x81\xe2", b"\xf0\x9f\x98", b"a\xc3", b"\xc3\x28",
        b" \xe2\x96\x81 ", b"\xe2\x96\x81", b"a\xe2\x96\x81 \xe2\x96\x81b",
    ]
    out = [s.encode("utf-8") for s in t] + raw
    alphabet = list("abcdefghij  \t\n") + ["é", "日", "ﬁ", "①", "▁", "\U0001F600", "́",
                                           " ", "Ａ", "ｶﾞ", "　", "​"]
    for _ in range(150):
        k = rnd.randrange(1, 60)
        out.append("".join(rnd.choice(alphabet) for _ in range(k)).encode("utf-8"))
    for _ in range(80):
        k = rnd.randrange(1, 40)
        out.append(bytes(rnd.getrandbits(8) for _ in range(k)))
    return out


def main():
    out_dir = os.path.join(HERE, "sp_bpe_8k")
    os.makedirs(out_dir, exist_ok=True)
    vocab = workload.make_vocabulary()
    corpus = workload.sentences(20000, seed=4321, vocabulary=vocab)
    model = io.BytesIO()
    spm.SentencePieceTrainer.train(
        sentence_iterator=iter(corpus), model_writer=model, model_type="bpe", vocab_size=8000,
        byte_fallback=True, character_coverage=1.0, num_threads=1, input_sentence_size=0,
        shuffle_input_sentence=False, minloglevel=2)
    with open(os.path.join(out_dir, "tokenizer.model"), "wb") as f:
        f.write(model.getvalue())
    with open(os.path.join(out_dir, "tokenizer_config.json"), "w") as f:
        json.dump({"tokenizer_class": "LlamaTokenizer", "add_bos_token": False, "add_eos_token": False}, f)
    sp = spm.SentencePieceProcessor(model_proto=model.getvalue())
    rnd = random.Random(99)
    texts = adversarial_texts(rnd) + [s.encode() for s in workload.sentences(150, seed=77, vocabulary=vocab)]
    texts.append(" ".join(workload.sentences(40, (30, 60), seed=5, vocabulary=vocab)).encode())
    gold = []
    for b in texts:
        ids = sp.EncodeAsIds(b)  # bytes go straight to the C++ Encode (no Python-side decoding)
        pieces = sp.EncodeAsPieces(b)
        norm = sp.Normalize(b)
        norm = norm.encode("utf-8") if isinstance(norm, str) else bytes(norm)
        gold.append({"text": b.hex(), "ids": ids, "n_pieces": len(pieces), "norm": norm.hex()})
    with open(os.path.join(HERE, "sp_bpe_8k_goldens.json"), "w") as f:
        json.dump({"sentencepiece_version": spm.__version__, "vocab_size": sp.GetPieceSize(), "cases": gold}, f,
                  separators=(",", ":"))
    print("model bytes", len(model.getvalue()), "cases", len(gold), "pieces", sp.GetPieceSize())


if __name__ == "__main__":
    main()
