"""GPU parity of the prefix index (write path + probe), GlobalKVCacheMgr::match and cache-aware routing
(csrc/prefix_index.cu) through the C-ABI against the CPU oracle, which keeps the reference's own
string-set containers.  Bit-exact: masks, scores, max_block_num, float routing scores; routing choice
must lie in the oracle's arg-max set (ties: SURVEY.md §7 "decision parity")."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N_INST = 24
NAMES = ["inst%02d" % i for i in range(N_INST)]


def _mk(oracle, capacity=1 << 16, block_size=128, seed=1024, names=NAMES):
    import xllm_service_b200 as x
    h = x.Ingest(block_size=block_size, xxh3_seed=seed, index_capacity=capacity)
    P = oracle.PrefixOracle(names, block_size, seed)
    return h, P


def _set_instances(h, P, rng, all_load=True, names=NAMES):
    for i, n in enumerate(names):
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


def _compare_requests(h, P, oracle, reqs, block_size=128, seed=1024, n_inst=N_INST):
    N_INST = n_inst  # noqa: N806
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


def test_all_64_instances(oracle):
    """The upper half of the instance masks (ids 32-63: the `lane + 32` score / candidate / arg-max paths of
    match_route_kernel) with every id in use, several prefill-side and decode-side candidates beyond id 31, ties
    across the two halves, and requests of more than 32 blocks (two probe waves, first miss in either)."""
    rng = np.random.default_rng(64)
    names = ["node-%02d" % i for i in range(64)]
    h, P = _mk(oracle, names=names)
    base = [rng.integers(0, 152000, size=128 * 70).astype(np.int32) for _ in range(4)]
    all_keys = [oracle.block_hash_chain(b) for b in base]
    reqs = []
    for i in range(150):
        b = base[int(rng.integers(0, len(base)))]
        cut = int(rng.integers(0, 71)) * 128
        reqs.append(np.concatenate([b[:cut], rng.integers(0, 152000, size=int(rng.integers(0, 300))).astype(np.int32)]))
    for window in range(8):
        _set_instances(h, P, rng, all_load=window % 2 == 0, names=names)
        if window == 5:     # only high ids usable: fallback and arg-max must come from lanes' second half
            for i, n in enumerate(names[:32]):
                P.set_instance(n, 1, False)
                h.set_instance(i, 1, False)
        for _ in range(60):
            i = int(rng.integers(0, 64)) if window % 3 else int(rng.integers(32, 64))
            kb = all_keys[int(rng.integers(0, len(base)))]
            upto = int(rng.integers(1, kb.shape[0] + 1))
            s = kb[:upto] if rng.random() < 0.7 else kb[rng.integers(0, kb.shape[0], size=int(rng.integers(0, 20)))]
            o = kb[rng.integers(0, kb.shape[0], size=int(rng.integers(0, 5)))]
            r = kb[rng.integers(0, kb.shape[0], size=int(rng.integers(0, 3)))]
            P.record(names[i], s, o, r)
            h.index_apply(i, s, o, r)
        P.upload()
        h.index_publish()
        assert h.index_size() == P.size()
        _compare_requests(h, P, oracle, reqs, n_inst=64)
    # at least some decisions really landed in the upper half
    lens = np.array([len(t) for t in reqs], dtype=np.int32)
    tok_start = np.zeros(len(reqs), np.int64)
    np.cumsum(lens[:-1], out=tok_start[1:])
    keys, key_start = h.hash_blocks(np.concatenate(reqs), tok_start, lens)
    match, routing = h.match_route(keys, key_start, (lens // 128).astype(np.int32))
    assert (routing["prefill_id"] >= 32).any() and (match["instances"] >> np.uint64(32)).any()
    assert (match["max_matched_block_num"] > 32).any()
    h.close()


def test_bulk_restore_beyond_64k_keys_and_failed_publish_does_not_poison(oracle):
    """A restarted master / a new replica lists the whole index from etcd and stages it with one put_bulk + one
    publish (host/index_snapshot.h apply_etcd_pairs): 200 K keys here (the packed op word used to hold a 16-bit
    payload index).  Then a publish that fails (over capacity) must leave the staging area empty so the next
    window's events go through."""
    import xllm_service_b200 as x
    rng = np.random.default_rng(11)
    n = 200_000
    h = x.Ingest(index_capacity=n + 1000)
    keys = rng.integers(0, 256, size=(n, 16)).astype(np.uint8)
    hbm = rng.integers(1, 2**63, size=n).astype(np.uint64)
    dram = rng.integers(0, 2**20, size=n).astype(np.uint64)
    ssd = np.zeros(n, np.uint64)
    h.index_put_bulk(keys, hbm, dram, ssd)
    h.index_publish()
    assert h.index_size() == n
    ek, eh, ed, es = h.index_export()
    order = np.lexsort(ek.T[::-1])
    want = np.lexsort(keys.T[::-1])
    assert (ek[order] == keys[want]).all() and (eh[order] == hbm[want]).all() and (ed[order] == dram[want]).all()
    for i in (0, 1, 65535, 65536, 65537, n - 1):
        assert h.index_get(keys[i]) == (True, [int(hbm[i]), int(dram[i]), 0])
    # over capacity -> error; afterwards the index still takes events
    h.index_apply(0, stored=rng.integers(0, 256, size=(5000, 16)).astype(np.uint8))
    with pytest.raises(x.IngestError) as ei:
        h.index_publish()
    assert ei.value.code == -6
    h.index_erase(keys[0])
    h.index_publish()
    assert h.index_get(keys[0])[0] is False
    h.close()


def test_churn_rebuilds_the_table_and_clear_instance(oracle):
    """Store / evict churn turns empty slots into tombstones; once live + tombstones pass 70 % of the slots publish
    re-inserts the live keys into a fresh table.  Content stays equal to the oracle's throughout, the tombstone
    count stays bounded, and xllm_index_clear_instance == `removed` events for every block of that instance."""
    rng = np.random.default_rng(5)
    cap = 2048                       # 4096 slots
    h, P = _mk(oracle, capacity=cap)
    _set_instances(h, P, rng)
    live = {}
    rebuilds_seen = 0
    for rnd in range(40):
        fresh = rng.integers(0, 256, size=(600, 16)).astype(np.uint8)
        inst = int(rng.integers(0, N_INST))
        P.record(NAMES[inst], stored=fresh)
        h.index_apply(inst, stored=fresh)
        for k in fresh:
            live[bytes(k)] = inst
        if len(live) > 1200:         # evict the oldest
            old = list(live.items())[:len(live) - 900]
            for inst_o in set(i for _, i in old):
                ks = np.array([np.frombuffer(k, np.uint8) for k, i in old if i == inst_o])
                P.record(NAMES[inst_o], removed=ks)
                h.index_apply(inst_o, removed=ks)
            for k, _ in old:
                del live[k]
        P.upload()
        h.index_publish()
        n_live, tombs, rebuilds = h.index_stats()
        assert n_live == P.size() == h.index_size() == len(live)
        assert (n_live + tombs) * 10 <= 4096 * 7
        rebuilds_seen = rebuilds
        for k in list(live)[::37]:
            assert h.index_get(np.frombuffer(k, np.uint8)) == P.get(np.frombuffer(k, np.uint8))
    assert rebuilds_seen >= 2
    # the instance holding most keys leaves the cluster
    inst = max(set(live.values()), key=lambda i: sum(1 for v in live.values() if v == i))
    mine = np.array([np.frombuffer(k, np.uint8) for k, i in live.items() if i == inst])
    P.record(NAMES[inst], removed=mine)
    P.upload()
    h.index_clear_instance(inst)
    assert h.index_size() == P.size()
    ek, eh, ed, es = h.index_export()
    assert not ((eh | ed | es) >> np.uint64(inst) & np.uint64(1)).any()
    for k in mine[::11]:
        assert h.index_get(k)[0] is False
    h.close()


def test_publish_from_a_clone_while_another_handle_matches(oracle):
    """Clones share one index (prefix_index.cuh reader / writer protocol): while one thread publishes windows that
    flip a key set between two complete states, another thread's match_route must see every request either entirely
    before or entirely after a publish — never a torn mixture (all 40 blocks of a prompt flip together)."""
    import ctypes
    import threading
    import xllm_service_b200 as x
    rng = np.random.default_rng(2)
    h = x.Ingest(index_capacity=1 << 14)
    c = ctypes.c_void_p()
    x._lib.check(h._L.xllm_ingest_clone(h._h, ctypes.byref(c)))
    w = x.Ingest.__new__(x.Ingest)
    w._L, w._h, w.block_size, w.seed = h._L, c, 128, 1024
    for i in range(4):
        h.set_instance(i, 1 if i % 2 == 0 else 2, True)
        h.set_load_metrics(i, 1, 0.5)
    toks = rng.integers(0, 1000, size=128 * 40).astype(np.int32)
    keys, key_start = h.hash_blocks(toks, np.zeros(1, np.int64), np.array([toks.size], np.int32))
    nreq = 256
    all_keys = np.tile(keys, (nreq, 1))
    ks = (np.arange(nreq) * 40).astype(np.int64)
    nb = np.full(nreq, 40, np.int32)
    stop = threading.Event()
    errors = []

    def writer():
        state = 0
        try:
            while not stop.is_set():
                state ^= 1
                # state 1: all 40 blocks on instance 0 ; state 0: all on instance 2 (one publish each)
                w.index_apply(0 if state else 2, stored=keys)
                w.index_apply(2 if state else 0, removed=keys)
                w.index_publish()
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    t = threading.Thread(target=writer)
    t.start()
    try:
        for _ in range(300):
            match, _ = h.match_route(all_keys, ks, nb)
            m = match["max_matched_block_num"]
            inst = match["instances"]
            assert ((m == 40) | (m == 0)).all(), np.unique(m)
            assert np.isin(inst[m == 40], [1, 4]).all(), np.unique(inst)
    finally:
        stop.set()
        t.join()
        w.close()
        h.close()
    assert not errors, errors


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
