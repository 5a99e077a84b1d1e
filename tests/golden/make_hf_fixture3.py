#!/usr/bin/env python3
"""The DeepSeek-V3 / R1 `tokenizer.json` layout on top of the trained vocabulary of hf_bpe_8k (make_hf_fixture.py):

  hf_deepseek_style/   normalizer Sequence[] (empty), pre_tokenizer Sequence[
                          Split(\\p{N}{1,3}, Isolated),
                          Split([一-龥぀-ゟ゠-ヿ]+, Isolated),
                          Split(<main regex: punct+ASCII letters | prefix? letters/marks | " ?" punct/symbols + CR/LF
                                 | \\s*[\\r\\n]+ | \\s+(?!\\S) | \\s+>, Isolated),
                          ByteLevel(add_prefix_space=false, use_regex=false)],
                       post_processor ByteLevel — the service names deepseek_v3 among its chat-parse families
                       (scheduler/xllm_chat_parse_bridge.cpp:49-78) and loads its tokenizer.json through FastTokenizer.

and freezes text -> ids goldens from upstream HF `tokenizers` (pip 0.22.2) through tokenizers_encode(text, 1)
(fast_tokenizer.cpp:20-30).  Outputs (committed): tests/golden/hf_deepseek_style/tokenizer.json,
tests/golden/hf_deepseek_goldens.json
"""
import json
import os
import random
import sys

from tokenizers import Tokenizer

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from xllm_service_b200 import workload  # noqa: E402

P_NUM = r"\p{N}{1,3}"
P_CJK = "[一-龥぀-ゟ゠-ヿ]+"
P_MAIN = ("[!\"#$%&'()*+,\\-./:;<=>?@\\[\\\\\\]^_`{|}~][A-Za-z]+|[^\r\n\\p{L}\\p{P}\\p{S}]?[\\p{L}\\p{M}]+|"
          " ?[\\p{P}\\p{S}]+[\r\n]*|\\s*[\r\n]+|\\s+(?!\\S)|\\s+")


def build(base):
    d = json.loads(json.dumps(base))
    d["normalizer"] = {"type": "Sequence", "normalizers": []}
    d["pre_tokenizer"] = {"type": "Sequence", "pretokenizers": [
        {"type": "Split", "pattern": {"Regex": P_NUM}, "behavior": "Isolated", "invert": False},
        {"type": "Split", "pattern": {"Regex": P_CJK}, "behavior": "Isolated", "invert": False},
        {"type": "Split", "pattern": {"Regex": P_MAIN}, "behavior": "Isolated", "invert": False},
        {"type": "ByteLevel", "add_prefix_space": False, "trim_offsets": True, "use_regex": False}]}
    d["post_processor"] = {"type": "ByteLevel", "add_prefix_space": True, "trim_offsets": False, "use_regex": True}
    os.makedirs(os.path.join(HERE, "hf_deepseek_style"), exist_ok=True)
    path = os.path.join(HERE, "hf_deepseek_style", "tokenizer.json")
    with open(path, "w", encoding="utf-8") as f:
        json.dump(d, f, ensure_ascii=False, separators=(",", ":"))
    return path


ALPHABET = list("abcdefghijXYZ  \t\n'.,!?012") + [
    "é", "日", "本", "語", "ひ", "ら", "カ", "ナ", "Σ", "١", " ", "　", "\U0001F600", "'s", "'RE", " '",
    "<|endoftext|>", "\r\n", " ", "_", "²", "½", "Ⅷ", "́", "̈", "​", "\x01", "\x7f", "­",
    "$", "+", "=", "<", ">", "^", "`", "|", "~", "€", "©", "→", "—", "“", "”", "。", "、", "・", "ー", "々", "〇",
    "\n\n", " \n", "\n ", "!\n", "12345", " 7", "x!y", " !", "\t!", "\r", "  ", ".com", "@user", "#tag", "(a)",
    "１２３", "٣٤٥٦", "ａｂｃ", "ß", "İ", "ǅ"]


def texts(rnd):
    t = ["", " ", "  ", "a", "Hello world", "Hello  world", "Hello world's  test\n\n 123 don't   x",
         "I'm you're we've they'll he'd it's can't 'tis O'Neil DON'T", "!!a x!a x !a  !a\t!a .!abc !.abc",
         "12345 1234567 1 12 123 1234 12日本 日本12 a1b2", "a \n \n b", "!\n\n x", "a   \n  b", "  x", "\t\tx",
         "x\n\ny", "'abc", " 's", "a\r\n\r\nb", "१२३४ a１２３４ ²³ ½x Ⅷz", "a \n", "a\n ", "a  ", "\n\n\nabc",
         "!!\n\n\n\n", "a  \n\n  \n  b", "x <|endoftext|> y", "a<|endoftext|>\nb", "tabs\tand\nnewlines\r\n\r\n  end  ",
         "numbers 3.14 1,000 1e10 x2y 2024-01-01", "(a) [b] {c} <d> a+b=c #tag @user $5 €9 ©x →y",
         "日本語のテキスト、カタカナとひらがな。漢字123abc", "中文字符和English混合text", "こんにちは世界！Hello",
         "éé äb ́x x́", "\x01\x02abc \x01 \x7fz a\x01b ​word​", "a­b",
         "foo.bar.baz www.example.com a@b.co x_y __init__ -v --flag a-b", "!abcé .xyzß ?Ωmega :abc1 ;abc日",
         "“quoted” ‘single’ —dash— …ellipsis", "x" * 300, "ab " * 200, "1234567890" * 30, "!?" * 100, "日本" * 150,
         " " * 40 + "z", "\n" * 20, " \n" * 20, "\n " * 20, "ｈｅｌｌｏ　ｗｏｒｌｄ", "ßtraße İstanbul ǅ"]
    out = list(t)
    for _ in range(400):
        out.append("".join(rnd.choice(ALPHABET) for _ in range(rnd.randrange(1, 80))))
    out += workload.sentences(40, (3, 80), seed=44)
    return out


def main():
    with open(os.path.join(HERE, "hf_bpe_8k", "tokenizer.json"), encoding="utf-8") as f:
        base = json.load(f)
    rnd = random.Random(5)
    cases = texts(rnd)
    tok = Tokenizer.from_file(build(base))
    gold = [{"text": s.encode("utf-8").hex(), "ids": tok.encode(s, add_special_tokens=True).ids} for s in cases]
    import tokenizers
    with open(os.path.join(HERE, "hf_deepseek_goldens.json"), "w") as f:
        json.dump({"tokenizers_version": tokenizers.__version__, "cases": gold}, f, separators=(",", ":"))
    print(len(gold), "cases")


if __name__ == "__main__":
    main()
