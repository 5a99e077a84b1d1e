"""SURVEY.md §8 (f)3 on the GPU: the published index leaves the device as etcd pairs (xllm_index_export ->
host/index_snapshot.h) and rebuilds an identical table on a replica whose instance ids differ
(xllm_index_put_bulk), then takes watch-style PUT / DELETE responses with the reference's puts-then-deletes order
(global_kvcache_mgr.cpp:47-51,133-175; etcd_client.cpp:122-137,174-198)."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(__file__)
ROOT = os.path.dirname(HERE)


def test_snapshot_restore_and_watch_events_cpp(tmp_path):
    exe = tmp_path / "index_snapshot_main"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "xllm_service_b200", "host"),
                           os.path.join(HERE, "cpp", "index_snapshot_main.cc"), "-o", str(exe),
                           "-L", os.path.join(ROOT, "xllm_service_b200"), "-lxllm_ingest",
                           "-Wl,-rpath," + os.path.join(ROOT, "xllm_service_b200")])
    p = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and p.stdout.startswith("OK "), (p.stdout, p.stderr)
    assert int(p.stdout.split()[1]) > 10000


def test_export_equals_the_oracle_map(oracle):
    """xllm_index_export returns exactly the reference's kvcache_infos_ after the same event history."""
    import xllm_service_b200 as x
    rng = np.random.default_rng(5)
    pool = rng.integers(0, 256, size=(3000, 16), dtype=np.uint8)
    h = x.Ingest(index_capacity=1 << 14)
    names = ["inst-%02d" % i for i in range(24)]
    P = oracle.PrefixOracle(names)
    try:
        for rnd in range(5):
            for _ in range(60):
                i = int(rng.integers(0, 24))
                st = pool[rng.integers(0, 3000, size=30)]
                off = pool[rng.integers(0, 3000, size=8)]
                rem = pool[rng.integers(0, 3000, size=6)]
                h.index_apply(i, st, off, rem)
                P.record(names[i], st, off, rem)
            h.index_publish()
            P.upload()
        keys, hbm, dram, ssd = h.index_export()
        assert keys.shape[0] == h.index_size() == P.size()
        assert len({bytes(k) for k in keys}) == keys.shape[0]
        for k, a, b, c in zip(keys, hbm, dram, ssd):
            found, want = P.get(bytes(k))
            assert found and [int(a), int(b), int(c)] == want
        # bulk put of the snapshot into a fresh handle gives the same table
        h2 = x.Ingest(index_capacity=1 << 14)
        try:
            h2.index_put_bulk(keys, hbm, dram, ssd)
            h2.index_publish()
            k2, a2, b2, c2 = h2.index_export()
            o1, o2 = np.lexsort(keys.T[::-1]), np.lexsort(k2.T[::-1])
            assert (keys[o1] == k2[o2]).all() and (hbm[o1] == a2[o2]).all() and (dram[o1] == b2[o2]).all() and (ssd[o1] == c2[o2]).all()
        finally:
            h2.close()
    finally:
        h.close()
