"""world_size-2 (and 4) gloo tests on CPU of the multi-rank HOST logic of the sharded index (the data path itself is
native NCCL and runs in tests/test_gpu_sharded.py on GPUs): the rendezvous helper that hands rank 0's id to every
rank, the owner function the ranks use to keep only their own slice of the event stream — union = everything,
pairwise disjoint, identical to the C-ABI's xllm_shard_owner — and bench.py's max-over-ranks reduction."""
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
        # rendezvous: one 128-byte blob minted on rank 0 reaches every rank unchanged (the NCCL id is replaced by a
        # stand-in here: minting a real one needs libnccl + a GPU)
        sharded.unique_id = lambda: bytes(range(128))
        uid = sharded.broadcast_unique_id()
        assert uid == bytes(range(128))
        # every rank sees the SAME event stream and keeps the keys it owns
        rng = np.random.default_rng(1234)
        keys = rng.integers(0, 256, size=(20000, 16), dtype=np.uint8)
        owner = sharded.owner_of_numpy(keys, world)
        mine = np.nonzero(owner == rank)[0]
        for i in mine[:200]:
            assert sharded.shard_owner(keys[i], world) == rank          # host C function == numpy restatement
        counts = torch.zeros(world, dtype=torch.int64)
        counts[rank] = mine.size
        dist.all_reduce(counts)
        assert int(counts.sum()) == keys.shape[0]                        # a partition: nothing lost, nothing doubled
        assert counts.min() > keys.shape[0] // world * 0.9               # and balanced (uniform hashes)
        digest = torch.zeros(world, dtype=torch.int64)
        digest[rank] = int(np.bitwise_xor.reduce(keys[mine].view("<u8")[:, 0].astype(np.uint64)) >> np.uint64(1))
        dist.all_reduce(digest)
        want = 0
        for r in range(world):
            want_r = np.bitwise_xor.reduce(keys[owner == r].view("<u8")[:, 0].astype(np.uint64)) >> np.uint64(1)
            assert int(digest[r]) == int(want_r)
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
def test_sharded_host_logic_gloo(world):
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
    assert sharded.owner_of_numpy(k, 1).tolist() == [0, 0, 0, 0]
    assert [sharded.shard_owner(x, 4) for x in k] == [0, 2, 1, 0]
    assert [sharded.shard_owner(x, 8) for x in k] == [0, 4, 2, 0]
