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


class SegmentBatch:
    """A batch of segmented requests in the layout xllm_ingest_batch_segments takes (include/xllm_ingest.h):
    text pieces back to back (text, offsets[n_pieces + 1]) + the segment table + the ready-made id spans."""

    def __init__(self, text, offsets, req_seg_start, seg_len, span_ids, offline, n_tokens):
        self.text, self.offsets = text, offsets
        self.req_seg_start, self.seg_len, self.span_ids = req_seg_start, seg_len, span_ids
        self.offline = offline            # bool [n_req]: best-effort requests (request/request.h:41)
        self.n_tokens = n_tokens          # int32 [n_req]: tokens each request must come out at

    @property
    def n(self):
        return self.req_seg_start.size - 1

    def select(self, rows):
        """The sub-batch of the given request rows (in that order), same layout."""
        rows = np.asarray(rows, dtype=np.int64)
        piece_of_seg = np.cumsum(self.seg_len < 0) - (self.seg_len < 0)          # piece index of a text segment
        span_of_seg = np.cumsum(np.maximum(self.seg_len, 0)) - np.maximum(self.seg_len, 0)
        texts, seg_len, spans, rss = [], [], [], [0]
        for r in rows:
            for s in range(self.req_seg_start[r], self.req_seg_start[r + 1]):
                ln = int(self.seg_len[s])
                seg_len.append(ln)
                if ln < 0:
                    p = piece_of_seg[s]
                    texts.append(self.text[self.offsets[p]:self.offsets[p + 1]])
                else:
                    spans.append(self.span_ids[span_of_seg[s]:span_of_seg[s] + ln])
            rss.append(len(seg_len))
        off = np.zeros(len(texts) + 1, np.int64)
        if texts:
            np.cumsum([t.size for t in texts], out=off[1:])
        return SegmentBatch(np.concatenate(texts) if texts else np.zeros(0, np.uint8), off,
                            np.asarray(rss, np.int32), np.asarray(seg_len, np.int32),
                            np.concatenate(spans).astype(np.int32) if spans else np.zeros(0, np.int32),
                            self.offline[rows], self.n_tokens[rows])


def make_c5_batch(n_requests, word_token_counts, seed=0, vocabulary=None, min_tokens=64, max_tokens=8192,
                  mm_frac=0.3, span_tokens=(256, 1024), max_spans=4, placeholder_ids=(7001, 7002, 7003),
                  offline_frac=0.3, shared_prefix=None):
    """BASELINE config 5 (SURVEY.md §8d C5 — synthetic, no reference semantics): the EPD multimodal mix.
      * request length log-uniform in [min_tokens, max_tokens] tokens;
      * mm_frac of the requests carry 1..max_spans runs of span_tokens[0]..span_tokens[1] repeated image-placeholder
        ids — already-tokenised spans that bypass BPE but are hashed and matched — between their text pieces;
      * offline_frac of the requests are offline (best effort; the batcher defers them behind online ones);
      * shared_prefix: optional dict(n_prefixes, frac, min_blocks, max_blocks, block_tokens): that fraction of the
        requests start their first text piece with one of the shared prefixes (Zipf-0.9 popularity).
    Every text piece is built from whole vocabulary words whose token counts (word_token_counts, measured with the
    tokenizer) add up exactly, so request r encodes to exactly n_tokens[r] tokens.  numpy only; deterministic."""
    vocab = vocabulary or make_vocabulary()
    V = len(vocab)
    wtc = np.asarray(word_token_counts, dtype=np.int64)
    rng = np.random.default_rng(seed)
    max_cnt = int(wtc.max())
    by_count = {c: np.nonzero(wtc == c)[0] for c in range(1, max_cnt + 1)}
    mean_tok = float(wtc[sample_word_ids(rng, 1, 4096, V)[0]].mean())
    pref_rows, pref_tok, pref_p = [], [], None
    if shared_prefix:
        sp = shared_prefix
        blocks = rng.integers(sp["min_blocks"], sp["max_blocks"] + 1, size=sp["n_prefixes"])
        for b in blocks:
            t = int(b) * sp["block_tokens"]
            pref_rows.append(_exact_row(rng, t, wtc, by_count, max_cnt, V, int(t / mean_tok * 1.2) + 64))
            pref_tok.append(t)
        pref_p = zipf_probs(sp["n_prefixes"])
    lengths = np.exp(rng.uniform(np.log(min_tokens), np.log(max_tokens + 1), size=n_requests)).astype(np.int64)
    lengths = np.clip(lengths, min_tokens, max_tokens)
    pieces, seg_len, spans, rss = [], [], [], [0]
    offline = rng.random(n_requests) < offline_frac
    prefix_id = np.full(n_requests, -1, np.int32)
    for r in range(n_requests):
        L = int(lengths[r])
        span_lens = []
        if rng.random() < mm_frac:
            if L < span_tokens[0] + 32:      # too short to hold a span: a multimodal request is at least one span long
                L = int(np.exp(rng.uniform(np.log(span_tokens[0] + 32), np.log(max_tokens + 1))))
                L = min(L, max_tokens)
                lengths[r] = L
            for _ in range(int(rng.integers(1, max_spans + 1))):
                ln = int(rng.integers(span_tokens[0], span_tokens[1] + 1))
                if sum(span_lens) + ln <= L - 16:          # keep at least a little text
                    span_lens.append(ln)
        text_tok = L - sum(span_lens)
        # text budget split over len(span_lens) + 1 pieces (a piece may come out empty and is then left out)
        cuts = np.sort(rng.integers(0, text_tok + 1, size=len(span_lens)))
        piece_tok = np.diff(np.concatenate([[0], cuts, [text_tok]]))
        lead = None
        if shared_prefix and rng.random() < shared_prefix["frac"]:
            j = int(rng.choice(len(pref_rows), p=pref_p))
            if pref_tok[j] <= piece_tok[0]:
                lead, prefix_id[r] = j, j
        for k, t in enumerate(piece_tok):
            t = int(t)
            if t > 0:
                words = []
                if k == 0 and lead is not None:
                    words.append(pref_rows[lead])
                    t -= pref_tok[lead]
                if t > 0:
                    words.append(_exact_row(rng, t, wtc, by_count, max_cnt, V, int(t / mean_tok * 1.2) + 64))
                ids = np.concatenate(words)
                pieces.append(b" ".join(vocab[i] for i in ids))
                seg_len.append(-1)
            if k < len(span_lens):
                spans.append(np.full(span_lens[k], placeholder_ids[int(rng.integers(0, len(placeholder_ids)))], np.int32))
                seg_len.append(span_lens[k])
        rss.append(len(seg_len))
    pb = pack_prompts(pieces)
    b = SegmentBatch(pb.text, pb.offsets, np.asarray(rss, np.int32), np.asarray(seg_len, np.int32),
                     np.concatenate(spans).astype(np.int32) if spans else np.zeros(0, np.int32), offline,
                     lengths.astype(np.int32))
    b.prefix_id = prefix_id
    b.prefix_tokens = np.array([pref_tok[j] if j >= 0 else 0 for j in prefix_id], np.int32)
    return b


def natural_corpus(max_bytes, exts=(".py", ".md", ".rst", ".txt", ".h", ".hpp", ".cuh"), min_file=4096, roots=None):
    """Real text already present in this image (SURVEY.md has no corpus; there is no network): the source and doc files
    under the Python site-packages tree, in sorted path order, UTF-8 only, concatenated until max_bytes.  Deterministic
    for a given image.  Returns one bytes object (files separated by a blank line)."""
    import os
    import sysconfig
    roots = roots or [sysconfig.get_paths()["purelib"]]
    files = []
    for root in roots:
        for dp, dn, fn in os.walk(root):
            dn.sort()
            for f in sorted(fn):
                if f.endswith(exts):
                    files.append(os.path.join(dp, f))
    out, total = [], 0
    for path in files:
        try:
            if os.path.getsize(path) < min_file:
                continue
            with open(path, "rb") as fh:
                data = fh.read()
            data.decode("utf-8")
        except (OSError, UnicodeDecodeError):
            continue
        if b"\x00" in data:
            continue
        out.append(data)
        total += len(data) + 2
        if total >= max_bytes:
            break
    return b"\n\n".join(out)[:max_bytes]


def cut_prompts(corpus: bytes, prompt_bytes=16384):
    """Slices of about prompt_bytes, cut at line ends (never inside a UTF-8 sequence) -> PromptBatch."""
    cuts, pos, n = [0], 0, len(corpus)
    while pos < n:
        end = min(pos + prompt_bytes, n)
        if end < n:
            nl = corpus.rfind(b"\n", pos + prompt_bytes // 2, end)
            if nl > pos:
                end = nl + 1
            else:
                while end < n and (corpus[end] & 0xC0) == 0x80:
                    end += 1
        cuts.append(end)
        pos = end
    return PromptBatch(np.frombuffer(corpus, dtype=np.uint8).copy(), np.asarray(cuts, dtype=np.int64))
