#!/usr/bin/env python3
"""BASELINE config 4: the hash-range-sharded prefix index over N GPUs — native NCCL exchange behind the C-ABI
(csrc/shard_exchange.cu).  Also the worker of tests/test_gpu_sharded.py.
Run:  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/sharded_check.py
Checks, on every rank, that the sharded path returns exactly what a replicated (full) index on the same GPU returns
for the same requests — device-pointer call, host-pointer call, an empty batch on one rank, a forced overflow round,
the whole xllm_ingest_batch pipeline — and on rank 0 that both equal the CPU oracle on a sample; then times the
round on the device (max over ranks) and prints one JSON line."""
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
from xllm_service_b200 import _lib, sharded, workload  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--requests", type=int, default=16384)
ap.add_argument("--blocks", type=int, default=32)
ap.add_argument("--index-keys", type=int, default=1 << 20)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--oracle-sample", type=int, default=64)
ap.add_argument("--no-pipeline", action="store_true")
a = ap.parse_args()

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")                # rendezvous only: the data path's NCCL communicator is the library's
dev = torch.device("cuda", local)
n, nb = a.requests, a.blocks
MODEL = os.path.join(ROOT, "tests", "golden", "sp_bpe_8k")
NAMES = ["inst%02d" % i for i in range(64)]

full = x.Ingest(tokenizer_path=MODEL, device=local, index_capacity=a.index_keys + (1 << 16))   # replica: the whole index
part = sharded.create_sharded(tokenizer_path=MODEL, device=local, index_capacity=a.index_keys // world * 2 + (1 << 16),
                              max_batch=n, max_tokens=nb * 128)                           # this rank's hash range
assert part.shard_last_stats()["bucket_capacity"] == n * nb * 3 // (2 * world) + 1024

rng = np.random.default_rng(7)                 # identical on every rank: the global event stream
idx = rng.integers(0, 256, size=(a.index_keys, 16), dtype=np.uint8)
inst = rng.integers(0, 64, size=a.index_keys)
loads = []
for i in range(64):
    k = idx[inst == i]
    off = k[rng.random(k.shape[0]) < 0.15]
    w, u = int(rng.integers(0, 16)), float(np.float32(rng.random()))
    loads.append((w, u))
    for h in (full, part):                     # every rank is given every event; the sharded handle keeps its own
        h.index_apply(i, k)
        h.index_apply(i, None, off)
        h.set_instance(i, 2 if i % 2 else 1, True)
        h.set_load_metrics(i, w, u)
full.index_publish()
part.index_publish()
sizes = torch.tensor([part.index_size()])
dist.all_reduce(sizes)
assert int(sizes.item()) == full.index_size() == a.index_keys, (int(sizes.item()), full.index_size())
own = sharded.owner_of_numpy(idx, world)
assert part.index_size() == int((own == rank).sum())
assert part.index_get(idx[np.nonzero(own == rank)[0][0]])[0] and not part.index_get(idx[np.nonzero(own != rank)[0][0]])[0]

r2 = np.random.default_rng(100 + rank)         # this rank's requests: prefixes of index keys, then misses
keys = r2.integers(0, 256, size=(n, nb, 16), dtype=np.uint8)
hit_len = r2.integers(0, nb + 1, size=n)
pick = r2.integers(0, a.index_keys, size=(n, nb))
m = np.arange(nb)[None, :] < hit_len[:, None]
keys[m] = idx[pick[m]]
n_blocks = r2.integers(0, nb + 1, size=n).astype(np.int32)     # ragged: not every row uses all its keys
n_blocks[: n // 2] = nb
d_keys = torch.from_numpy(keys.reshape(-1, 16)).to(dev)
d_ks = torch.arange(n, device=dev, dtype=torch.int64) * nb
d_nb = torch.from_numpy(n_blocks).to(dev)


def outs(k=n):
    return torch.zeros((k, 400), dtype=torch.uint8, device=dev), torch.zeros((k, 20), dtype=torch.uint8, device=dev)


# ---- 1. device-pointer call: sharded == replicated, bit for bit
m_full, r_full = outs()
m_sh, r_sh = outs()
full.match_route_device(n, d_keys.data_ptr(), n * nb, d_ks.data_ptr(), d_nb.data_ptr(), m_full.data_ptr(),
                        r_full.data_ptr(), None)
part.match_route_device(n, d_keys.data_ptr(), n * nb, d_ks.data_ptr(), d_nb.data_ptr(), m_sh.data_ptr(),
                        r_sh.data_ptr(), None)
torch.cuda.synchronize()
assert torch.equal(m_full, m_sh) and torch.equal(r_full, r_sh), "sharded result differs from the replicated index"
mm = m_full.cpu().numpy().view(_lib.MATCH_DTYPE)[:, 0]
assert (mm["max_matched_block_num"] == np.minimum(hit_len, n_blocks)).all()
assert part.shard_last_stats()["overflow_rounds"] == 0

# ---- 2. host-pointer call; rank 1 brings an EMPTY batch (the round is collective and must not dead-lock)
k2 = 0 if rank == 1 else 257
mh, rh = part.match_route(keys[:k2].reshape(-1, 16), np.arange(k2, dtype=np.int64) * nb, n_blocks[:k2])
mf, rf = full.match_route(keys[:k2].reshape(-1, 16), np.arange(k2, dtype=np.int64) * nb, n_blocks[:k2])
assert mh.tobytes() == mf.tobytes() and rh.tobytes() == rf.tobytes()

# ---- 3. overflow: every key of the batch owned by ONE rank (the same key repeated) -> bucket > capacity -> every
# rank repeats the round with the same larger capacity, results still exact
hot = idx[np.nonzero(own == (world - 1))[0][:4]]
kk = np.tile(hot[r2.integers(0, 4, size=n * nb)].reshape(n, nb, 16), 1)
dk = torch.from_numpy(kk.reshape(-1, 16)).to(dev)
d_nb_full = torch.full((n,), nb, dtype=torch.int32, device=dev)
mo_f, ro_f = outs()
mo_s, ro_s = outs()
full.match_route_device(n, dk.data_ptr(), n * nb, d_ks.data_ptr(), d_nb_full.data_ptr(), mo_f.data_ptr(), ro_f.data_ptr(), None)
part.match_route_device(n, dk.data_ptr(), n * nb, d_ks.data_ptr(), d_nb_full.data_ptr(), mo_s.data_ptr(), ro_s.data_ptr(), None)
torch.cuda.synchronize()
assert torch.equal(mo_f, mo_s) and torch.equal(ro_f, ro_s), "overflow round: sharded differs from replicated"
st = part.shard_last_stats()
assert st["overflow_rounds"] >= 1 and st["bucket_capacity"] >= n * nb, st
# ... and the ordinary batch still works afterwards
part.match_route_device(n, d_keys.data_ptr(), n * nb, d_ks.data_ptr(), d_nb.data_ptr(), m_sh.data_ptr(),
                        r_sh.data_ptr(), None)
torch.cuda.synchronize()
assert torch.equal(m_full, m_sh) and torch.equal(r_full, r_sh)

# ---- 4. the CPU oracle on a sample (rank 0): the index content the sample touches, in the oracle's containers
if rank == 0 and a.oracle_sample:
    from oracle import oracle as o
    P = o.PrefixOracle(NAMES)
    for i, nme in enumerate(NAMES):
        P.set_instance(nme, 2 if i % 2 else 1)
        P.set_load(nme, *loads[i])
    S = min(a.oracle_sample, n)
    touched = np.unique(pick[:S][m[:S]])
    for j in touched:
        f, masks = full.index_get(idx[j])
        P.put(idx[j], *[[NAMES[b] for b in range(64) if (mk >> b) & 1] for mk in masks])
    mmf = m_sh.cpu().numpy().view(_lib.MATCH_DTYPE)[:, 0]
    rrf = r_sh.cpu().numpy().view(_lib.ROUTING_DTYPE)[:, 0]
    for r in range(S):
        ks = keys[r, :n_blocks[r]]
        # oracle match over precomputed keys: walk them as GlobalKVCacheMgr::match does
        matched, hbm = 0, np.zeros(64, np.uint32)
        for b, kq in enumerate(ks):
            f, masks = P.get(kq)
            if not f or not any(masks):
                break
            matched = b + 1
            for bit in range(64):
                if (masks[0] >> bit) & 1:
                    hbm[bit] = b + 1
        assert mmf["max_matched_block_num"][r] == matched and mmf["max_block_num"][r] == n_blocks[r]
        assert (mmf["hbm"][r] == hbm).all()

# ---- 5. the whole pipeline: text in, ids / keys / match / routing out, sharded == replicated
if not a.no_pipeline:
    texts = [s.encode() for s in workload.sentences(600 + 37 * rank, (200, 700), seed=50 + rank)]
    b = workload.pack_prompts(texts)
    T = 1024
    ref = full.ingest_batch(b.text, b.offsets, T)
    # make the index know some of these prompts' prefixes (same events on every rank)
    allk = [None] * world
    mine = [ref["keys"][r, : ref["n_ids"][r] // 128][: 1 + r % 5] for r in range(0, b.n, 3)]
    dist.all_gather_object(allk, np.concatenate(mine))
    for src, kk2 in enumerate(allk):
        for h in (full, part):
            h.index_apply(src, kk2)
    full.index_publish()
    part.index_publish()
    part.set_pipeline(97, 1 << 20)
    full.set_pipeline(97, 1 << 20)
    ref = full.ingest_batch(b.text, b.offsets, T)
    got = part.ingest_batch(b.text, b.offsets, T)
    for f in ("n_ids", "status", "keys"):
        assert (ref[f] == got[f]).all(), f
    valid = np.arange(T)[None, :] < ref["n_ids"][:, None]          # ids past n_ids are not part of the result
    assert (ref["ids"][valid] == got["ids"][valid]).all()
    assert ref["match"].tobytes() == got["match"].tobytes() and ref["routing"].tobytes() == got["routing"].tobytes()
    assert (got["match"]["max_matched_block_num"] > 0).any()

# ---- timing: the round on the device, max over ranks
times, parts = [], []
for it in range(a.iters + 2):
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    part.match_route_device(n, d_keys.data_ptr(), n * nb, d_ks.data_ptr(), d_nb.data_ptr(), m_sh.data_ptr(),
                            r_sh.data_ptr(), torch.cuda.current_stream().cuda_stream or None)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if it >= 2:
        times.append(float(t.item()))
        parts.append(part.shard_last_stats())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    full.match_route_device(n, d_keys.data_ptr(), n * nb, d_ks.data_ptr(), d_nb.data_ptr(), m_full.data_ptr(),
                            r_full.data_ptr(), torch.cuda.current_stream().cuda_stream or None)
e1.record()
torch.cuda.synchronize()
if rank == 0:
    ms = sorted(times)[len(times) // 2]
    med = {k: sorted(p[k] for p in parts)[len(parts) // 2] for k in ("bucket_ms", "exchange_out_ms", "probe_ms",
                                                                    "exchange_back_ms", "score_ms")}
    print(json.dumps({"check": "sharded == replicated == oracle(sample)", "n_gpus": world, "requests_per_gpu": n,
                      "blocks": nb, "index_keys": a.index_keys, "keys_in_rank0_shard": part.index_size(),
                      "round_ms_max_over_ranks": round(ms, 4), "match_req_per_s": round(world * n / ms * 1e3),
                      "round_parts_ms_rank0": {k: round(v, 4) for k, v in med.items()},
                      "replicated_fused_match_ms": round(e0.elapsed_time(e1) / a.iters, 4),
                      "tuple_bytes_out_per_rank": int(n_blocks.sum()) * 24, "mask_bytes_back_per_rank": int(n_blocks.sum()) * 24,
                      "bucket_capacity": parts[-1]["bucket_capacity"]}))
part.close()
full.close()
dist.barrier()
dist.destroy_process_group()
