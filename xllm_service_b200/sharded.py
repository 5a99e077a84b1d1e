"""Hash-range-sharded prefix index across the box's GPUs (BASELINE config 4) — rendezvous only.

The data path is native: libxllm_ingest.so buckets the block keys by owner on the device, exchanges 24-byte
(hash128, request, block) tuples with one grouped ncclSend/ncclRecv round, probes on the owning GPU, sends the tier
masks back with a second round and runs the first-miss scan + cache-aware routing on the origin GPU
(csrc/shard_exchange.cu), all behind the ordinary C-ABI calls (xllm_match_route, xllm_ingest_batch) of a handle
created with `shard_world / shard_rank / nccl_unique_id`.  What is left for the host is to hand one rank's
ncclUniqueId to the others; a C++ service does that over whatever channel it has (etcd, the launcher), this helper
does it over an existing torch.distributed process group (any backend: gloo works, no tensors on the GPU needed).
"""
import ctypes

import numpy as np

from . import _lib
from .ingest import Ingest


def owner_of_numpy(keys_u8: np.ndarray, world: int) -> np.ndarray:
    """keys_u8 [n,16] uint8 -> owner rank per key: the top log2(world) bits of the key's low64 (the same function the
    device bucketing kernel and xllm_shard_owner apply)."""
    assert world >= 1 and world & (world - 1) == 0, "world size must be a power of two"
    bits = world.bit_length() - 1
    low = np.ascontiguousarray(keys_u8.reshape(-1, 16)[:, :8]).view("<u8")[:, 0]
    return (low >> np.uint64(64 - bits)).astype(np.int64) if bits else np.zeros(low.shape[0], np.int64)


def shard_owner(key16, world: int) -> int:
    """xllm_shard_owner through the C-ABI (host-only: works without a GPU)."""
    k = np.ascontiguousarray(np.frombuffer(bytes(key16), dtype=np.uint8))
    rc = _lib.lib().xllm_shard_owner(ctypes.c_void_p(k.ctypes.data), world)
    if rc < 0:
        _lib.check(rc)
    return rc


def unique_id() -> bytes:
    """xllm_shard_unique_id: 128 bytes to be passed to every rank's create call."""
    buf = ctypes.create_string_buffer(128)
    _lib.check(_lib.lib().xllm_shard_unique_id(buf))
    return buf.raw


def broadcast_unique_id(group=None) -> bytes:
    """Rank 0 of the (already initialised) torch.distributed group mints the id, everybody receives it."""
    import torch.distributed as dist
    box = [unique_id() if dist.get_rank(group) == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    return box[0]


def create_sharded(group=None, **ingest_kwargs) -> Ingest:
    """Collective: every rank of the group calls this with the same configuration (device = its own GPU)."""
    import torch.distributed as dist
    uid = broadcast_unique_id(group)
    return Ingest(shard_world=dist.get_world_size(group), shard_rank=dist.get_rank(group), nccl_unique_id=uid,
                  **ingest_kwargs)
