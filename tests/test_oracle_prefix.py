"""Hand-built cases for the CPU oracle of match / index updates / cache-aware routing
(oracle/prefix_oracle.cc).  The reference has no tests for these; every expected
value below is derived by hand from the cited reference lines."""
import numpy as np
import pytest

NAMES = ["p0", "p1", "p2", "d0", "d1"]


@pytest.fixture()
def P(oracle):
    p = oracle.PrefixOracle(NAMES)
    for n, t in zip(NAMES, [p.PREFILL, p.DEFAULT, p.MIX, p.DECODE, p.DECODE]):
        p.set_instance(n, t)
    return p


def _keys(oracle, toks):
    return oracle.block_hash_chain(toks, 128, 1024)


def test_first_miss_and_tail_tokens(oracle, P):
    toks = np.arange(128 * 6 + 77, dtype=np.int32)  # 6 full blocks + 77 tail tokens (ignored, :76)
    k = _keys(oracle, toks)
    P.record("p0", stored=k[[0, 1, 2, 4, 5]])  # block 3 missing
    P.upload()
    m = P.match(toks)
    assert m["max_block_num"] == 6
    assert m["max_matched_block_num"] == 3  # stops at the first miss (:127-129)
    assert m["hbm"].tolist() == [3, 0, 0, 0, 0]
    assert m["instances"] == 0b00001
    # fewer tokens than one block: untouched OverlapScores (:77-79)
    m = P.match(toks[:100])
    assert m["max_block_num"] == 0 and m["instances"] == 0


def test_non_contiguous_presence_scores_last_index(oracle, P):
    toks = np.arange(128 * 5, dtype=np.int32)
    k = _keys(oracle, toks)
    P.record("p0", stored=k[:5])
    P.record("p1", stored=k[[0, 3]])  # present at blocks 0 and 3 only: score = 3 + 1 (set_score overwrites, :64-68)
    P.record("d0", stored=k[[1]])
    P.upload()
    m = P.match(toks)
    assert m["hbm"].tolist() == [5, 4, 0, 2, 0]
    assert m["max_matched_block_num"] == 5


def test_tier_moves_and_empty_value_is_a_miss(oracle, P):
    toks = np.arange(128 * 3, dtype=np.int32)
    k = _keys(oracle, toks)
    P.record("p0", stored=k[:3])
    P.upload()
    # offload: HBM -> DRAM (:201-203); second offload: not in HBM -> DRAM erased, SSD inserted (:204-207)
    P.record("p0", offload=k[1:2])
    P.upload()
    assert P.get(k[1]) == (True, [0, 1, 0])
    m = P.match(toks)
    assert m["hbm"].tolist()[0] == 3 and m["dram"].tolist()[0] == 2  # hbm score overwritten at block 2
    P.record("p0", offload=k[1:2])
    P.upload()
    assert P.get(k[1]) == (True, [0, 0, 1])
    # offload of an instance that holds nothing: goes straight to SSD (the reference's else-branch quirk)
    P.record("p1", offload=k[0:1])
    P.upload()
    assert P.get(k[0]) == (True, [1, 0, 2])
    # removed from all tiers; the entry becomes empty and upload erases it (:236-238) -> miss at block 1
    P.record("p0", removed=k[1:2])
    P.upload()
    assert P.get(k[1])[0] is False
    m = P.match(toks)
    assert m["max_matched_block_num"] == 1
    # offload / removed of unknown keys are skipped (:195-197,213-215): no entry is created
    other = _keys(oracle, np.arange(1000, 1128, dtype=np.int32))
    n0 = P.size()
    P.record("p0", offload=other, removed=other)
    P.upload()
    assert P.size() == n0


def test_staging_is_invisible_until_upload_and_keeps_empties(oracle, P):
    toks = np.arange(128, dtype=np.int32)
    k = _keys(oracle, toks)
    P.record("p0", stored=k)
    assert P.match(toks)["max_matched_block_num"] == 0  # staged only (master flush every 3 s)
    # removed then offload inside one window: the staged (empty) copy is still there, so the offload
    # is NOT skipped and lands in SSD (:198-207)
    P.record("p0", removed=k)
    P.record("p0", offload=k)
    P.upload()
    assert P.get(k[0]) == (True, [0, 0, 1])


def test_replica_put_delete(oracle, P):
    toks = np.arange(256, dtype=np.int32)
    k = _keys(oracle, toks)
    P.put(k[0], hbm=["p0", "d0"], dram=["p1"])
    P.put(k[1], ssd=["d1"])
    m = P.match(toks)
    assert m["hbm"].tolist() == [1, 0, 0, 1, 0] and m["dram"].tolist() == [0, 1, 0, 0, 0]
    assert m["ssd"].tolist() == [0, 0, 0, 0, 2] and m["instances"] == 0b11011
    P.delete(k[0])
    assert P.match(toks)["max_matched_block_num"] == 0


def test_cost_function_integer_divisions(oracle, P):
    toks = np.arange(128 * 4, dtype=np.int32)
    k = _keys(oracle, toks)
    P.record("p0", stored=k[:4])   # full hit  -> first term 1
    P.record("p1", stored=k[:3])   # 3/4       -> first term 0 (unsigned integer division, :73)
    P.record("d0", stored=k[:4])
    P.record("d1", stored=k[:1])
    P.upload()
    P.set_load("p0", 10, 0.75)
    P.set_load("p1", 4, 0.25)
    P.set_load("d0", 2, 0.5)
    P.set_load("d1", 1, 0.125)
    r = P.route(toks)
    assert r["ok"]
    # prefill: p0 = 1 - 0.75 - 10/10 = -0.75 ; p1 = 0 - 0.25 - 4/10(=0) = -0.25  -> p1
    assert r["prefill_id"] == 1 and r["prefill_score"] == -0.25
    # decode: d0 = 1 - 0.5 - 2/2 = -0.5 ; d1 = 0 - 0.125 - 1/2(=0) = -0.125 -> d1
    assert r["decode_id"] == 4 and r["decode_score"] == -0.125


def test_fallback_and_no_node(oracle, P):
    toks = np.arange(128 * 2, dtype=np.int32)
    # nothing matched: both sides fall back to the least gpu_cache_usage_perc (< 1, strict) (:314-358)
    P.set_load("p0", 5, 0.9)
    P.set_load("p2", 9, 0.4)
    P.set_load("d0", 1, 1.0)   # usage 1 is never < 1
    P.set_load("d1", 1, 0.99)
    r = P.route(toks)
    assert r["ok"] and r["prefill_id"] == 2 and r["decode_id"] == 4
    # fallback candidates keep max_waiting == 0 => third term 0: score = 0 - usage
    assert r["prefill_score"] == pytest.approx(-0.4) and r["decode_score"] == pytest.approx(-0.99)
    # unschedulable / metric-less instances are ignored; no prefill-side candidate -> false (:35-38)
    P.set_instance("p0", P.PREFILL, False)
    P.set_instance("p2", P.MIX, False)
    assert P.route(toks)["ok"] is False
    # an empty token list skips match entirely (:25-30) and still routes by fallback
    P.set_instance("p2", P.MIX, True)
    assert P.route(np.zeros(0, np.int32))["prefill_id"] == 2


def test_matched_but_unschedulable_is_not_a_candidate(oracle, P):
    toks = np.arange(128, dtype=np.int32)
    k = _keys(oracle, toks)
    P.record("p0", stored=k)
    P.record("d0", stored=k)
    P.upload()
    P.set_load("p0", 0, 0.1)
    P.set_load("p1", 0, 0.6)
    P.set_load("d1", 0, 0.2)   # d0 matched but has no metrics -> skipped (:291-295); decode falls back to d1
    P.set_instance("p0", P.PREFILL, False)
    r = P.route(toks)
    assert r["ok"] and r["prefill_id"] == 1 and r["decode_id"] == 4


def test_ties_report_argmax_set(oracle, P):
    toks = np.arange(128, dtype=np.int32)
    k = _keys(oracle, toks)
    P.record("p0", stored=k)
    P.record("p1", stored=k)
    P.upload()
    P.set_load("p0", 2, 0.5)
    P.set_load("p1", 2, 0.5)
    r = P.route(toks)
    assert r["prefill_argmax"] == 0b00011 and r["prefill_id"] in (0, 1)
    # scores that never beat MIN_SCORE (-2.0, strict >) leave the name empty
    P.set_load("p0", 2, 1.5)
    P.set_load("p1", 2, 1.0)
    r = P.route(toks)   # 1 - 1.5 - 1 = -1.5 ; 1 - 1.0 - 1 = -1.0
    assert r["prefill_id"] == 1
    P.set_load("p0", 2, 2.5)
    P.set_load("p1", 2, 2.0)
    r = P.route(toks)   # -2.5 and -2.0: neither is > -2.0
    assert r["ok"] and r["prefill_id"] == -1 and r["prefill_argmax"] == 0
