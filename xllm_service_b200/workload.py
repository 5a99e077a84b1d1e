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


def _materialise(words, row_nwords, word_len, word_off, flat):
    """torch: flat word ids of consecutive rows + words per row -> (text uint8, row byte lengths);
    words joined by single spaces."""
    import torch
    dev = words.device
    wl = word_len[words]
    n_rows = row_nwords.numel()
    row_first = torch.cumsum(row_nwords, 0) - row_nwords
    row_of = torch.repeat_interleave(torch.arange(n_rows, device=dev), row_nwords)
    step = wl + 1                                            # word + one separator
    excl = torch.cumsum(step, 0) - step                      # byte offset if every word kept its separator
    lens = torch.zeros(n_rows, dtype=torch.int64, device=dev)
    lens.index_add_(0, row_of, step)
    lens = torch.where(row_nwords > 0, lens - 1, lens)       # each row drops its last separator
    row_base = torch.cumsum(lens, 0) - lens
    first_excl = torch.zeros(n_rows, dtype=torch.int64, device=dev)
    nz = row_nwords > 0
    first_excl[nz] = excl[row_first[nz]]
    start = row_base[row_of] + (excl - first_excl[row_of])
    lexcl = torch.cumsum(wl, 0) - wl
    within = torch.arange(int(wl.sum()), device=dev) - torch.repeat_interleave(lexcl, wl)
    text = torch.full((int(lens.sum()),), 32, dtype=torch.uint8, device=dev)
    text[torch.repeat_interleave(start, wl) + within] = flat[torch.repeat_interleave(word_off[words], wl) + within]
    return text, lens


def make_prompts_exact_tokens(n_prompts, n_tokens, word_token_counts, seed=0, vocabulary=None,
                              shared_prefix=None, chunk=2048, device="cpu"):
    """Build n_prompts prompts that each encode to EXACTLY n_tokens tokens under a tokenizer for
    which words tokenize independently (SentencePiece-BPE with a whitespace-split vocabulary):
    word_token_counts[i] = number of tokens of the word "▁" + vocabulary[i].  Words are drawn
    Zipf(0.9); the last word of a prompt is re-drawn from the words with exactly the remaining count.
    Array work runs in torch on `device` ("cuda" for the 1 GB batches of bench.py); the result is
    deterministic in (seed, device type).

    shared_prefix: optional dict(n_prefixes, frac, min_blocks, max_blocks, block_tokens) — the
    BASELINE config-3 workload: `frac` of the prompts start with one of n_prefixes shared
    prefixes (popularity Zipf(0.9)) whose length is uniform in [min_blocks, max_blocks] KV blocks.

    Returns (PromptBatch, meta) with meta = dict(prefix_id int32[n] (-1 = none),
    prefix_blocks int32[n], n_prefixes).
    """
    import torch
    vocab = vocabulary or make_vocabulary()
    V = len(vocab)
    wtc_np = np.asarray(word_token_counts, dtype=np.int64)
    assert wtc_np.size == V and wtc_np.min() >= 1
    rng = np.random.default_rng(seed)
    max_cnt = int(wtc_np.max())
    by_count = {c: np.nonzero(wtc_np == c)[0] for c in range(1, max_cnt + 1)}
    assert all(by_count[c].size > 0 for c in range(1, max_cnt)), "need a word for every remainder"
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed) * 7919 + 17)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    wtc = T(wtc_np)
    word_len = T(np.array([len(w) for w in vocab], dtype=np.int64))
    flat = T(np.frombuffer(b"".join(vocab), dtype=np.uint8).copy())
    word_off = torch.cumsum(word_len, 0) - word_len
    zipf_cdf = T(np.cumsum(zipf_probs(V)))
    by_count_t = {c: T(v) for c, v in by_count.items()}

    est_words = int(n_tokens / wtc_np[sample_word_ids(rng, 1, 4096, V)[0]].mean() * 1.1) + 64
    n_pref = 0
    pref_blk_np = np.zeros(0, np.int64)
    if shared_prefix:
        sp = shared_prefix
        n_pref = sp["n_prefixes"]
        pref_blk_np = rng.integers(sp["min_blocks"], sp["max_blocks"] + 1, size=n_pref)
        pref_rows = [_exact_row(rng, int(b) * sp["block_tokens"], wtc_np, by_count, max_cnt, V, est_words)
                     for b in pref_blk_np]
        pref_tok = T(pref_blk_np * sp["block_tokens"])
        pref_cdf = T(np.cumsum(zipf_probs(n_pref)))
        pref_len = T(np.array([len(p) for p in pref_rows], dtype=np.int64))
        pref_flat = T(np.concatenate(pref_rows).astype(np.int64))
        pref_off = torch.cumsum(pref_len, 0) - pref_len
    prefix_id = np.full(n_prompts, -1, dtype=np.int32)
    texts, lens_all = [], []
    for c0 in range(0, n_prompts, chunk):
        m = min(chunk, n_prompts - c0)
        pid = torch.full((m,), -1, dtype=torch.int64, device=dev)
        if shared_prefix:
            use = torch.rand(m, generator=gen, device=dev) < shared_prefix["frac"]
            pick = torch.clamp(torch.searchsorted(pref_cdf, torch.rand(m, generator=gen, device=dev,
                                                                       dtype=torch.float64)), max=n_pref - 1)
            pid = torch.where(use & (pref_tok[pick] <= n_tokens), pick, pid)
        has = pid >= 0
        budget = torch.full((m,), n_tokens, dtype=torch.int64, device=dev)
        plen = torch.zeros(m, dtype=torch.int64, device=dev)
        if shared_prefix:
            safe = torch.clamp(pid, min=0)
            budget = torch.where(has, budget - pref_tok[safe], budget)
            plen = torch.where(has, pref_len[safe], plen)
        u = torch.rand((m, est_words), generator=gen, device=dev, dtype=torch.float64)
        ids = torch.clamp(torch.searchsorted(zipf_cdf, u), max=V - 1)
        csum = torch.cumsum(wtc[ids], 1)
        k = (csum <= budget[:, None]).sum(1)                                # body words that fit entirely
        assert bool((k < est_words).all())
        used = torch.where(k > 0, csum.gather(1, torch.clamp(k - 1, min=0)[:, None])[:, 0], torch.zeros_like(k))
        rem = budget - used                                                  # 0 .. max_cnt - 1
        tail = torch.full((m,), -1, dtype=torch.int64, device=dev)
        for c in range(1, max_cnt):
            sel = rem == c
            cnt = int(sel.sum())
            if cnt:
                tail[sel] = by_count_t[c][torch.randint(0, by_count_t[c].numel(), (cnt,), generator=gen, device=dev)]
        assert bool(((rem == 0) | (tail >= 0)).all())
        has_tail = tail >= 0
        nw = plen + k + has_tail.to(torch.int64)
        width = int(nw.max())
        M = torch.full((m, width), -1, dtype=torch.int64, device=dev)
        if shared_prefix and bool(has.any()):                                # prefix words
            hp = plen[has]
            rr = torch.repeat_interleave(torch.nonzero(has)[:, 0], hp)
            cc = torch.arange(int(hp.sum()), device=dev) - torch.repeat_interleave(torch.cumsum(hp, 0) - hp, hp)
            M[rr, cc] = pref_flat[torch.repeat_interleave(pref_off[pid[has]], hp) + cc]
        body = torch.arange(est_words, device=dev)[None, :] < k[:, None]
        rc = torch.nonzero(body)
        M[rc[:, 0], plen[rc[:, 0]] + rc[:, 1]] = ids[rc[:, 0], rc[:, 1]]
        tr = torch.nonzero(has_tail)[:, 0]
        M[tr, (plen + k)[tr]] = tail[tr]
        words = M[M >= 0]
        text, lens = _materialise(words, nw, word_len, word_off, flat)
        texts.append(text.cpu().numpy())
        lens_all.append(lens.cpu().numpy())
        prefix_id[c0:c0 + m] = pid.cpu().numpy()
    lens = np.concatenate(lens_all) if lens_all else np.zeros(0, np.int64)
    offsets = np.zeros(n_prompts + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    text = np.concatenate(texts) if texts else np.zeros(0, np.uint8)
    pb = np.zeros(n_prompts, dtype=np.int32)
    if shared_prefix:
        pb[prefix_id >= 0] = pref_blk_np[prefix_id[prefix_id >= 0]]
    meta = {"prefix_id": prefix_id, "prefix_blocks": pb, "n_prefixes": n_pref}
    return PromptBatch(text, offsets), meta
