#!/usr/bin/env python3
"""BASELINE config 4: hash-range-sharded prefix index over N GPUs with one NCCL all-to-all each way per batch.
Run:  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/sharded_check.py
Checks that the sharded path returns exactly what a replicated (full) index returns for the same
requests, then times probe+exchange+score on the device (max over ranks)."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xllm_service_b200 as x  # noqa: E402
from xllm_service_b200 import _lib, sharded  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--requests", type=int, default=16384)
ap.add_argument("--blocks", type=int, default=32)
ap.add_argument("--index-keys", type=int, default=1 << 20)
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
n, nb = a.requests, a.blocks

full = x.Ingest(device=local, index_capacity=a.index_keys + 4096)       # replica: the whole index
part = x.Ingest(device=local, index_capacity=a.index_keys // world * 2 + 4096)  # this rank's hash range only
sh = sharded.ShardedIndex(part)

rng = np.random.default_rng(7)                 # identical on every rank: the global event stream
idx = rng.integers(0, 256, size=(a.index_keys, 16), dtype=np.uint8)
inst = rng.integers(0, 64, size=a.index_keys)
for i in range(64):
    k = idx[inst == i]
    full.index_apply(i, k)
    sh.apply(i, k)
    off = k[rng.random(k.shape[0]) < 0.15]
    full.index_apply(i, None, off)
    sh.apply(i, None, off)
    w, u = int(rng.integers(0, 16)), float(np.float32(rng.random()))
    for h in (full, part):
        h.set_instance(i, 2 if i % 2 else 1, True)
        h.set_load_metrics(i, w, u)
full.index_publish()
sh.publish()
sizes = torch.tensor([part.index_size()], device=dev)
dist.all_reduce(sizes)
assert int(sizes.item()) == full.index_size(), (int(sizes.item()), full.index_size())

r2 = np.random.default_rng(100 + rank)         # this rank's requests: prefixes of index keys, then misses
keys = np.zeros((n, nb, 16), np.uint8)
hit_len = r2.integers(0, nb + 1, size=n)
pick = r2.integers(0, a.index_keys, size=(n, nb))
keys[:] = r2.integers(0, 256, size=(n, nb, 16), dtype=np.uint8)
m = np.arange(nb)[None, :] < hit_len[:, None]
keys[m] = idx[pick[m]]
d_keys = torch.from_numpy(keys.reshape(-1, 16)).to(dev)
d_ks = torch.arange(n, device=dev, dtype=torch.int64) * nb
d_nb = torch.full((n,), nb, dtype=torch.int32, device=dev)


def outs():
    return torch.zeros((n, 400), dtype=torch.uint8, device=dev), torch.zeros((n, 20), dtype=torch.uint8, device=dev)


m_full, r_full = outs()
m_sh, r_sh = outs()
s = torch.cuda.current_stream().cuda_stream or None
full.match_route_device(n, d_keys.data_ptr(), n * nb, d_ks.data_ptr(), d_nb.data_ptr(), m_full.data_ptr(),
                        r_full.data_ptr(), s)
sh.match_route(d_keys, d_ks, d_nb, n, m_sh, r_sh)
torch.cuda.synchronize()
assert torch.equal(m_full, m_sh) and torch.equal(r_full, r_sh), "sharded result differs from the replicated index"
mm = m_full.cpu().numpy().view(_lib.MATCH_DTYPE)[:, 0]
assert (mm["max_matched_block_num"] == hit_len).all()

times = []
for it in range(a.iters + 2):
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sh.match_route(d_keys, d_ks, d_nb, n, m_sh, r_sh)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if it >= 2:
        times.append(float(t.item()))
if rank == 0:
    ms = sorted(times)[len(times) // 2]
    print(json.dumps({"check": "sharded == replicated", "n_gpus": world, "requests_per_gpu": n, "blocks": nb,
                      "index_keys": a.index_keys, "keys_per_rank_index": part.index_size(),
                      "match_ms_max_over_ranks": ms, "match_req_per_s": world * n / ms * 1e3,
                      "alltoall_bytes_per_rank_each_way": [n * nb * 16, n * nb * 24]}))
dist.destroy_process_group()
