"""Synthetic workload generator for tests and bench.py: numpy only.

Vocabulary: 20 000 pseudo-words of 2-9 lowercase letters (PRNG seed 1234); text: words drawn
Zipf(s = 0.9), space separated.  Everything is deterministic in its seed arguments.
"""
import numpy as np

N_WORDS = 20000
VOCAB_SEED = 1234
ZIPF_S = 0.9


def make_vocabulary(n_words=N_WORDS, seed=VOCAB_SEED):
    """Returns a list of n_words distinct byte strings of 2-9 lowercase letters."""
    rng = np.random.default_rng(seed)
    words, seen = [], set()
    while len(words) < n_words:
        ln = int(rng.integers(2, 10))
        w = bytes(rng.integers(97, 123, size=ln, dtype=np.uint8))
        if w not in seen:
            seen.add(w)
            words.append(w)
    return words


def zipf_probs(n, s=ZIPF_S):
    p = 1.0 / np.arange(1, n + 1, dtype=np.float64) ** s
    return p / p.sum()


def sample_word_ids(rng, n_rows, n_cols, n_words=N_WORDS, s=ZIPF_S):
    """int32 [n_rows, n_cols] of word indices drawn Zipf(s) (inverse-CDF sampling)."""
    cdf = np.cumsum(zipf_probs(n_words, s))
    u = rng.random((n_rows, n_cols))
    return np.minimum(np.searchsorted(cdf, u), n_words - 1).astype(np.int32)


def sentences(n_sentences, words_per_sentence=(8, 40), seed=0, vocabulary=None):
    """Training / test corpus: list of str."""
    vocab = vocabulary or make_vocabulary()
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_sentences):
        k = int(rng.integers(words_per_sentence[0], words_per_sentence[1] + 1))
        ids = sample_word_ids(rng, 1, k, len(vocab))[0]
        out.append(b" ".join(vocab[i] for i in ids).decode("ascii"))
    return out


class PromptBatch:
    """A batch of prompts in the layout the C-ABI takes: one contiguous uint8 text buffer +
    int64 offsets [n+1]."""

    def __init__(self, text, offsets):
        self.text = text
        self.offsets = offsets

    @property
    def n(self):
        return self.offsets.size - 1

    def prompt(self, i):
        return self.text[self.offsets[i]:self.offsets[i + 1]].tobytes()


def pack_prompts(prompts):
    """list[bytes|str] -> PromptBatch."""
    bs = [p.encode("utf-8") if isinstance(p, str) else bytes(p) for p in prompts]
    offsets = np.zeros(len(bs) + 1, dtype=np.int64)
    if bs:
        np.cumsum([len(b) for b in bs], out=offsets[1:])
    text = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, np.uint8)
    return PromptBatch(text, offsets)


def _fill_exact(rng, rem, by_count, max_cnt):
    """Word ids whose token counts sum to exactly `rem`."""
    tail = []
    while rem > 0:
        c = min(rem, max_cnt)
        while by_count[c].size == 0:
            c -= 1
        tail.append(int(rng.choice(by_count[c])))
        rem -= c
    return tail


def _exact_row(rng, n_tokens, wtc, by_count, max_cnt, V, est_words):
    ids = sample_word_ids(rng, 1, max(est_words, 64), V)[0]
    csum = np.cumsum(wtc[ids])
    k = int(np.searchsorted(csum, n_tokens, side="right"))
    rem = n_tokens - (int(csum[k - 1]) if k else 0)
    return np.concatenate([ids[:k], np.asarray(_fill_exact(rng, rem, by_count, max_cnt), dtype=np.int32)]).astype(
        np.int32)


def make_prompts_exact_tokens(n_prompts, n_tokens, word_token_counts, seed=0, vocabulary=None,
                              shared_prefix=None):
    """Build n_prompts prompts that each encode to EXACTLY n_tokens tokens under a tokenizer for
    which words tokenize independently (SentencePiece-BPE with a whitespace-split vocabulary):
    word_token_counts[i] = number of tokens of the word "▁" + vocabulary[i].  Words are drawn
    Zipf(0.9); the tail of a prompt is re-drawn from the words with exactly the remaining count.

    shared_prefix: optional dict(n_prefixes, frac, min_blocks, max_blocks, block_tokens) — the
    BASELINE config-3 workload: `frac` of the prompts start with one of n_prefixes shared
    prefixes (popularity Zipf(0.9)) whose length is uniform in [min_blocks, max_blocks] KV blocks.

    Returns (PromptBatch, list of per-prompt int32 word-id arrays).
    """
    vocab = vocabulary or make_vocabulary()
    V = len(vocab)
    wtc = np.asarray(word_token_counts, dtype=np.int64)
    assert wtc.size == V and wtc.min() >= 1
    rng = np.random.default_rng(seed)
    max_cnt = int(wtc.max())
    by_count = {c: np.nonzero(wtc == c)[0] for c in range(1, max_cnt + 1)}
    assert by_count[1].size > 0, "need at least one single-token word to fill exactly"
    word_len = np.array([len(w) for w in vocab], dtype=np.int64)
    flat = np.frombuffer(b"".join(vocab), dtype=np.uint8)
    word_off = np.zeros(V + 1, dtype=np.int64)
    np.cumsum(word_len, out=word_off[1:])

    est_words = int(n_tokens / wtc[sample_word_ids(rng, 1, 4096, V)[0]].mean() * 1.15) + 8
    prefixes = None
    if shared_prefix:
        sp = shared_prefix
        prefixes = [_exact_row(rng, int(rng.integers(sp["min_blocks"], sp["max_blocks"] + 1)) * sp["block_tokens"],
                               wtc, by_count, max_cnt, V, est_words) for _ in range(sp["n_prefixes"])]
        pref_tok = [int(wtc[p].sum()) for p in prefixes]
        pref_p = zipf_probs(sp["n_prefixes"])
    rows = []
    chunk = 2048
    for c0 in range(0, n_prompts, chunk):
        m = min(chunk, n_prompts - c0)
        ids = sample_word_ids(rng, m, est_words, V)
        csum = np.cumsum(wtc[ids], axis=1)
        for i in range(m):
            budget = n_tokens
            head = None
            if prefixes is not None and rng.random() < shared_prefix["frac"]:
                j = int(rng.choice(len(prefixes), p=pref_p))
                if pref_tok[j] <= n_tokens:
                    head = prefixes[j]
                    budget = n_tokens - pref_tok[j]
            k = int(np.searchsorted(csum[i], budget, side="right"))  # words that fit entirely
            rem = budget - (int(csum[i, k - 1]) if k else 0)
            parts = [ids[i, :k]]
            if rem:
                parts.append(np.asarray(_fill_exact(rng, rem, by_count, max_cnt), dtype=np.int32))
            if head is not None:
                parts.insert(0, head)
            rows.append(np.concatenate(parts).astype(np.int32))
    return rows_to_batch(rows, vocab, word_len, word_off, flat), rows


def rows_to_batch(rows, vocab, word_len=None, word_off=None, flat=None):
    """Materialise word-id rows as text: words joined by single spaces."""
    if word_len is None:
        word_len = np.array([len(w) for w in vocab], dtype=np.int64)
        flat = np.frombuffer(b"".join(vocab), dtype=np.uint8)
        word_off = np.zeros(len(vocab) + 1, dtype=np.int64)
        np.cumsum(word_len, out=word_off[1:])
    n = len(rows)
    lens = np.array([int(word_len[r].sum()) + max(len(r) - 1, 0) for r in rows], dtype=np.int64)
    offsets = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    text = np.full(int(offsets[-1]), 32, dtype=np.uint8)
    for i, r in enumerate(rows):
        if len(r) == 0:
            continue
        wl = word_len[r]
        tot = int(wl.sum())
        excl = np.concatenate([[0], np.cumsum(wl[:-1])])          # exclusive prefix of word lengths
        dst0 = offsets[i] + excl + np.arange(len(r))               # +1 space per preceding word
        ar = np.arange(tot) - np.repeat(excl, wl)                  # byte index inside its word
        text[np.repeat(dst0, wl) + ar] = flat[np.repeat(word_off[r], wl) + ar]
    return PromptBatch(text, offsets)
