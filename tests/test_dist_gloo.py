"""world_size-2 (and 4) gloo tests on CPU of the multi-rank host logic: the hash-range owner function, the
bucket / all-to-all / probe / all-to-all / unbucket exchange of xllm_service_b200/sharded.py with a
dictionary standing in for the device probe, and bench.py's max-over-ranks reduction."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from xllm_service_b200 import sharded
        rng = np.random.default_rng(1234)            # every rank builds the SAME global index description
        n_index = 5000
        idx_keys = rng.integers(0, 256, size=(n_index, 16), dtype=np.uint8)
        idx_masks = rng.integers(1, 2**62, size=(n_index, 3), dtype=np.int64)
        owner = sharded.owner_of_numpy(idx_keys, world)
        assert set(np.unique(owner)) <= set(range(world))
        local = {idx_keys[i].tobytes(): idx_masks[i] for i in np.nonzero(owner == rank)[0]}

        def probe(keys):
            out = torch.zeros((keys.shape[0], 3), dtype=torch.int64)
            for j in range(keys.shape[0]):
                kb = keys[j].numpy().tobytes()
                # the exchange must only ever ask the owner
                assert sharded.owner_of_numpy(keys[j].numpy()[None, :], world)[0] == rank
                if kb in local:
                    out[j] = torch.from_numpy(local[kb])
            return out

        ex = sharded.ShardedExchange(probe)
        r2 = np.random.default_rng(99 + rank)        # each rank asks for its own batch of keys
        pick = r2.integers(0, n_index, size=700 + 13 * rank)
        q_keys = idx_keys[pick].copy()
        miss = r2.random(q_keys.shape[0]) < 0.3
        q_keys[miss] = r2.integers(0, 256, size=(int(miss.sum()), 16), dtype=np.uint8)
        got = ex.lookup(torch.from_numpy(q_keys)).numpy()
        want = np.where(miss[:, None], 0, idx_masks[pick])
        assert (got == want).all()
        # empty batch on one rank must not dead-lock the others
        got0 = ex.lookup(torch.zeros((0 if rank == 0 else 5, 16), dtype=torch.uint8))
        assert got0.shape[0] == (0 if rank == 0 else 5)
        # bench.py's max-over-ranks timing reduction
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == world
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, "FAIL: %r" % (e,)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_exchange_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(world)], res


def test_owner_function_is_top_bits():
    from xllm_service_b200 import sharded
    k = np.zeros((4, 16), np.uint8)
    k[1, 7] = 0x80          # low64 top bit set
    k[2, 7] = 0x40
    k[3, 15] = 0xFF         # high64 must not matter
    assert sharded.owner_of_numpy(k, 2).tolist() == [0, 1, 0, 0]
    assert sharded.owner_of_numpy(k, 4).tolist() == [0, 2, 1, 0]
    low = torch.from_numpy(np.ascontiguousarray(k[:, :8]).view("<i8")[:, 0].copy())
    assert sharded.owner_of(low, 4).tolist() == [0, 2, 1, 0]
    assert sharded.owner_of(low, 1).tolist() == [0, 0, 0, 0]
