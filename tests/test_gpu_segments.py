"""BASELINE config 5 on the GPU: requests made of text pieces and ready-made token-id spans
(xllm_ingest_batch_segments) against the CPU oracle.  The oracle side restates the reference's own append semantics —
every text piece is one Tokenizer::encode call whose ids are appended to Request::token_ids (scheduler.cpp:128-132,
sentencepiece_tokenizer.cpp:122-126), id spans are appended as they are — then hashes / matches / routes the
concatenation with the usual oracle (GlobalKVCacheMgr::match + CacheAwareRouting restatement).  Variable lengths
(log-uniform 64-8192 tokens), 30 % of requests with 1-4 placeholder-id spans, ragged chunk boundaries, empty pieces,
span-only and text-only requests, truncation, both tokenizer families."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(__file__)
SP_DIR = os.path.join(HERE, "golden", "sp_bpe_8k")
HF_DIR = os.path.join(HERE, "golden", "hf_llama3_style")
NAMES = ["inst%02d" % i for i in range(16)]


def _oracle_ids(encode_piece, b, r):
    """concatenation of per-piece encodes and id spans of request r"""
    piece_of_seg = np.cumsum(b.seg_len < 0) - (b.seg_len < 0)
    span_of_seg = np.cumsum(np.maximum(b.seg_len, 0)) - np.maximum(b.seg_len, 0)
    out = []
    for s in range(b.req_seg_start[r], b.req_seg_start[r + 1]):
        ln = int(b.seg_len[s])
        if ln < 0:
            p = piece_of_seg[s]
            out.extend(encode_piece(b.text[b.offsets[p]:b.offsets[p + 1]].tobytes()))
        else:
            out.extend(b.span_ids[span_of_seg[s]:span_of_seg[s] + ln].tolist())
    return np.asarray(out, np.int32)


def _word_counts(h, vocab):
    from xllm_service_b200 import workload
    wb = workload.pack_prompts(vocab)
    _, n, st = h.encode_batch(wb.text, wb.offsets, 32)
    assert (st == 0).all()
    return n


def test_c5_mix_matches_oracle(oracle):
    import xllm_service_b200 as x
    from xllm_service_b200 import workload
    rng = np.random.default_rng(5)
    h = x.Ingest(tokenizer_path=SP_DIR, index_capacity=1 << 15)
    h.set_pipeline(41, 1 << 20)          # small odd chunks: piece / span / segment offsets cross many chunk borders
    sp = oracle.SentencePieceOracle(SP_DIR)
    P = oracle.PrefixOracle(NAMES)
    vocab = workload.make_vocabulary()
    wcnt = _word_counts(h, vocab)
    b = workload.make_c5_batch(260, wcnt, seed=9, shared_prefix=dict(n_prefixes=6, frac=0.6, min_blocks=1, max_blocks=5,
                                                                     block_tokens=128))
    assert (b.seg_len >= 0).sum() > 40 and b.offline.sum() > 40
    for i, n in enumerate(NAMES):
        t = 2 if i % 2 else 1
        w, u = int(rng.integers(0, 9)), float(np.float32(rng.random()))
        P.set_instance(n, t)
        P.set_load(n, w, u)
        h.set_instance(i, t)
        h.set_load_metrics(i, w, u)
    T = 8192
    want = [_oracle_ids(lambda t: sp.encode(t).tolist(), b, r) for r in range(b.n)]
    assert [w.size for w in want] == b.n_tokens.tolist()           # the generator's exact-length contract
    # index: prefixes of some requests (these include placeholder-span blocks: they are hashed like any other id)
    for r in range(0, b.n, 2):
        keys = oracle.block_hash_chain(want[r])
        i = int(rng.integers(0, len(NAMES)))
        k = keys[:int(rng.integers(0, keys.shape[0] + 1))]
        P.record(NAMES[i], k)
        h.index_apply(i, k)
    P.upload()
    h.index_publish()
    out = h.ingest_batch_segments(b, T)
    assert (out["status"] == 0).all()
    n_span_blocks_matched = 0
    for r in range(b.n):
        n = want[r].size
        assert out["n_ids"][r] == n, r
        assert (out["ids"][r, :n] == want[r]).all(), r
        keys = oracle.block_hash_chain(want[r])
        assert (out["keys"][r, :keys.shape[0]] == keys).all(), r
        assert not out["keys"][r, keys.shape[0]:].any(), r
        m = P.match(want[r])
        assert out["match"]["max_block_num"][r] == m["max_block_num"], r
        assert out["match"]["max_matched_block_num"][r] == m["max_matched_block_num"], r
        assert out["match"]["hbm"][r][:len(NAMES)].tolist() == m["hbm"].tolist(), r
        ro = P.route(want[r])
        assert bool(out["routing"]["ok"][r]) == ro["ok"]
        assert out["routing"]["prefill_score"][r] == np.float32(ro["prefill_score"]), r
        assert (ro["prefill_argmax"] >> int(out["routing"]["prefill_id"][r])) & 1, r
        # a matched block that lies inside a placeholder span: pre-tokenised ids took part in the match
        first_span = next((s for s in range(b.req_seg_start[r], b.req_seg_start[r + 1]) if b.seg_len[s] >= 0), None)
        if first_span is not None and m["max_matched_block_num"] * 128 >= n // 2:
            n_span_blocks_matched += 1
    assert n_span_blocks_matched > 0
    # the online / offline split is a host-side selection of rows: both halves give the same per-request results
    on = b.select(np.nonzero(~b.offline)[0])
    o2 = h.ingest_batch_segments(on, T, want_match=False)
    rows = np.nonzero(~b.offline)[0]
    assert (o2["n_ids"] == out["n_ids"][rows]).all() and (o2["keys"] == out["keys"][rows]).all()
    h.close()


def test_edge_shapes(oracle):
    """span-only request, text-only request, empty text piece, empty request, adjacent spans, truncated row, and a
    piece the device refuses (NFC) failing only its own request — on the HF backend with template ids per piece."""
    import xllm_service_b200 as x
    from xllm_service_b200 import workload
    H = oracle.HfBpeOracle(HF_DIR)

    def enc(t):
        return H.prefix_ids + H.encode(t).tolist() + H.suffix_ids

    reqs = [
        [np.arange(300, 340, dtype=np.int32)],                                   # span only
        [b"hello world, plain text"],                                            # text only
        [b"", np.full(7, 42, np.int32), b""],                                    # empty pieces still add template ids
        [],                                                                      # no segments at all
        [np.full(5, 1, np.int32), np.full(6, 2, np.int32), b" tail", np.full(3, 3, np.int32)],
        [b"a b c d e f g h i j k l m n o p", np.arange(100, dtype=np.int32)],    # will be truncated at stride 64
        ["café naïve".encode(), np.full(4, 9, np.int32), "déjà vu".encode()],
    ]
    pieces, seg_len, spans, rss = [], [], [], [0]
    for r in reqs:
        for s in r:
            if isinstance(s, bytes):
                pieces.append(s)
                seg_len.append(-1)
            else:
                spans.append(s)
                seg_len.append(s.size)
        rss.append(len(seg_len))
    pb = workload.pack_prompts(pieces)
    b = workload.SegmentBatch(pb.text, pb.offsets, np.asarray(rss, np.int32), np.asarray(seg_len, np.int32),
                              np.concatenate(spans).astype(np.int32), np.zeros(len(reqs), bool),
                              np.zeros(len(reqs), np.int32))
    h = x.Ingest(tokenizer_path=HF_DIR)
    out = h.ingest_batch_segments(b, 64, want_match=False)
    for r in range(len(reqs)):
        want = _oracle_ids(enc, b, r)
        if r == 5:
            assert out["status"][r] == 1 and out["n_ids"][r] == want.size and want.size > 64
            assert (out["ids"][r] == want[:64]).all()
        else:
            assert out["status"][r] == 0, r
            assert out["n_ids"][r] == want.size and (out["ids"][r, :want.size] == want).all(), r
    h.close()
    # Qwen2 layout (normalizer NFC): a piece that is not NFC fails its request with XLLM_ERR_UNSUPPORTED, others pass
    hq = x.Ingest(tokenizer_path=os.path.join(HERE, "golden", "hf_qwen2_style"))
    pieces = [b"fine", "é decomposed".encode(), b"also fine"]
    pb = workload.pack_prompts(pieces)
    b = workload.SegmentBatch(pb.text, pb.offsets, np.array([0, 1, 3, 4], np.int32), np.array([-1, 5, -1, -1], np.int32),
                              np.arange(5, dtype=np.int32), np.zeros(3, bool), np.zeros(3, np.int32))
    out = hq.ingest_batch_segments(b, 64, want_match=False)
    assert out["status"].tolist() == [0, -5, 0] and out["n_ids"][1] == 0
    hq.close()
