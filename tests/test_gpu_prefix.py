"""GPU parity of the prefix index (write path + probe), GlobalKVCacheMgr::match and cache-aware routing
(csrc/prefix_index.cu) through the C-ABI against the CPU oracle, which keeps the reference's own
string-set containers.  Bit-exact: masks, scores, max_block_num, float routing scores; routing choice
must lie in the oracle's arg-max set (ties: SURVEY.md §7 "decision parity")."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N_INST = 24
NAMES = ["inst%02d" % i for i in range(N_INST)]


def _mk(oracle, capacity=1 << 16, block_size=128, seed=1024):
    import xllm_service_b200 as x
    h = x.Ingest(block_size=block_size, xxh3_seed=seed, index_capacity=capacity)
    P = oracle.PrefixOracle(NAMES, block_size, seed)
    return h, P


def _set_instances(h, P, rng, all_load=True):
    for i, n in enumerate(NAMES):
        t = int(rng.choice([0, 1, 2, 2, 3]))
        sched = bool(rng.random() > 0.1)
        P.set_instance(n, t, sched)
        h.set_instance(i, t, sched)
        if all_load or rng.random() > 0.2:
            w = int(rng.integers(0, 6))
            u = float(np.float32(rng.choice([0.0, 0.125, 0.25, 0.5, 0.75, 0.99, 1.0, 1.5]) if rng.random() < 0.5
                                 else rng.random()))
            P.set_load(n, w, u)
            h.set_load_metrics(i, w, u)


def _compare_requests(h, P, oracle, reqs, block_size=128, seed=1024):
    lens = np.array([len(t) for t in reqs], dtype=np.int32)
    tok_start = np.zeros(len(reqs), np.int64)
    np.cumsum(lens[:-1], out=tok_start[1:])
    toks = np.concatenate(reqs).astype(np.int32) if len(reqs) else np.zeros(0, np.int32)
    keys, key_start = h.hash_blocks(toks, tok_start, lens)
    n_blocks = (lens // block_size).astype(np.int32)
    match, routing = h.match_route(keys, key_start, n_blocks)
    for r, t in enumerate(reqs):
        m = P.match(t)
        assert match["max_block_num"][r] == m["max_block_num"], r
        assert match["max_matched_block_num"][r] == m["max_matched_block_num"], r
        assert match["instances"][r] == m["instances"], r
        assert match["hbm"][r][:N_INST].tolist() == m["hbm"].tolist(), r
        assert match["dram"][r][:N_INST].tolist() == m["dram"].tolist(), r
        assert match["ssd"][r][:N_INST].tolist() == m["ssd"].tolist(), r
        ro = P.route(t)
        assert bool(routing["ok"][r]) == ro["ok"], r
        if not ro["ok"]:
            continue
        assert routing["prefill_score"][r] == np.float32(ro["prefill_score"]), r
        pid = int(routing["prefill_id"][r])
        if ro["prefill_argmax"] == 0:
            assert pid == -1
        else:
            assert pid >= 0 and (ro["prefill_argmax"] >> pid) & 1, (r, pid, ro)
        did = int(routing["decode_id"][r])
        if ro["decode_id"] == -1 and ro["decode_argmax"] == 0:
            assert did == -1
        else:
            assert routing["decode_score"][r] == np.float32(ro["decode_score"]), r
            assert did >= 0 and (ro["decode_argmax"] >> did) & 1, (r, did, ro)


def test_random_events_match_route(oracle):
    rng = np.random.default_rng(7)
    h, P = _mk(oracle)
    _set_instances(h, P, rng)
    # a family of prompts sharing prefixes
    base = [rng.integers(0, 152000, size=128 * 40).astype(np.int32) for _ in range(6)]
    reqs = []
    for i in range(120):
        b = base[int(rng.integers(0, len(base)))]
        cut = int(rng.integers(0, 41)) * 128
        tail = rng.integers(0, 152000, size=int(rng.integers(0, 700))).astype(np.int32)
        reqs.append(np.concatenate([b[:cut], tail]))
    reqs += [np.zeros(0, np.int32), base[0][:127], base[0][:128]]
    all_keys = [oracle.block_hash_chain(b) for b in base]
    for window in range(12):
        for _ in range(30):
            i = int(rng.integers(0, N_INST))
            kb = all_keys[int(rng.integers(0, len(base)))]
            pick = lambda n: kb[rng.integers(0, kb.shape[0], size=int(rng.integers(0, n)))]  # noqa: E731
            s, o, r = pick(30), pick(8), pick(6)
            P.record(NAMES[i], s, o, r)
            h.index_apply(i, s, o, r)
        P.upload()
        h.index_publish()
        assert h.index_size() == P.size()
        for kb in all_keys:
            for k in kb[::7]:
                fo, mo = P.get(k)
                fg, mg = h.index_get(k)
                assert mo == mg and fo == fg
        if window % 3 == 2:
            _set_instances(h, P, rng, all_load=False)
        _compare_requests(h, P, oracle, reqs)
    h.close()


def test_replica_put_erase_and_capacity(oracle):
    import xllm_service_b200 as x
    rng = np.random.default_rng(3)
    h, P = _mk(oracle, capacity=64)
    _set_instances(h, P, rng)
    toks = rng.integers(0, 1000, size=128 * 20).astype(np.int32)
    keys = oracle.block_hash_chain(toks)
    for j, k in enumerate(keys):
        hb = [NAMES[i] for i in range(N_INST) if (j + i) % 5 == 0]
        dr = [NAMES[i] for i in range(N_INST) if (j * i) % 11 == 3]
        P.put(k, hbm=hb, dram=dr)
        h.index_put(k, sum(1 << NAMES.index(n) for n in hb), sum(1 << NAMES.index(n) for n in dr), 0)
    h.index_publish()
    _compare_requests(h, P, oracle, [toks, toks[:128 * 7 + 5]])
    P.delete(keys[4])
    h.index_erase(keys[4])
    h.index_publish()
    _compare_requests(h, P, oracle, [toks])
    assert h.index_size() == P.size()
    # tombstoned slot is reusable and the key can come back
    P.put(keys[4], hbm=[NAMES[1]])
    h.index_put(keys[4], 2, 0, 0)
    h.index_publish()
    _compare_requests(h, P, oracle, [toks])
    # over capacity: loud failure
    many = rng.integers(0, 256, size=(100, 16)).astype(np.uint8)
    h.index_apply(0, stored=many)
    with pytest.raises(x.IngestError):
        h.index_publish()
    h.close()


def test_no_index_is_an_error():
    import xllm_service_b200 as x
    h = x.Ingest()
    with pytest.raises(x.IngestError):
        h.index_publish()
    h.close()
