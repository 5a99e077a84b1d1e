"""Hash-range-sharded prefix index across the box's GPUs (BASELINE config 4).

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).  Every rank holds the slice of
the index whose keys fall in its hash range (owner = top log2(G) bits of the key's low64 — keys are
XXH3-128 outputs, so ranges are uniformly loaded).  Tokenize + block-hash stay data-parallel over
requests; per batch the ONLY data-path collective is one all-to-all that ships each rank's block keys
to their owners and one all-to-all that brings the three tier masks (24 B) per key back:

    keys[n,16] --bucket by owner--> all_to_all --> local probe --> all_to_all --> unbucket --> score + route

torch.distributed is plumbing here; the probe (xllm_index_probe_device) and the first-miss scan +
cache-aware routing (xllm_score_route_device) are this repo's CUDA kernels.  The same class runs on
CPU tensors with the gloo backend and a pluggable `probe_fn`, which is how the exchange logic is
tested without GPUs (tests/test_dist_gloo.py).
"""
import numpy as np
import torch
import torch.distributed as dist


def owner_of(low64: torch.Tensor, world: int) -> torch.Tensor:
    """low64: int64 tensor holding the keys' low 64 bits (bit pattern).  owner = low64 >> (64 - log2 G)."""
    assert world & (world - 1) == 0, "world size must be a power of two"
    bits = world.bit_length() - 1
    if bits == 0:
        return torch.zeros_like(low64)
    # logical shift of the unsigned bit pattern: arithmetic shift then mask
    return (low64 >> (64 - bits)) & (world - 1)


def owner_of_numpy(keys_u8: np.ndarray, world: int) -> np.ndarray:
    """keys_u8 [n,16] uint8 -> owner rank per key (host side: routes KvCacheEvents to the owning rank)."""
    bits = world.bit_length() - 1
    low = np.ascontiguousarray(keys_u8.reshape(-1, 16)[:, :8]).view("<u8")[:, 0]
    return (low >> np.uint64(64 - bits)).astype(np.int64) if bits else np.zeros(low.shape[0], np.int64)


class ShardedExchange:
    """keys -> owners -> probe -> masks back, for the calling rank's batch."""

    def __init__(self, probe_fn, group=None):
        """probe_fn(keys uint8 [k,16] tensor) -> int64 [k,3] tensor of (hbm, dram, ssd) masks on this shard."""
        self.probe_fn = probe_fn
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def lookup(self, keys: torch.Tensor) -> torch.Tensor:
        """keys: uint8 [n,16] on this rank's device.  Returns int64 [n,3] masks in the caller's key order."""
        n = keys.shape[0]
        dev = keys.device
        low = keys[:, :8].contiguous().view(torch.int64)[:, 0]
        owner = owner_of(low, self.world)
        order = torch.argsort(owner, stable=True)
        send_counts = torch.bincount(owner, minlength=self.world)
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts, group=self.group)
        sc, rc = send_counts.tolist(), recv_counts.tolist()
        send_keys = keys[order].contiguous()
        recv_keys = torch.empty((sum(rc), 16), dtype=torch.uint8, device=dev)
        dist.all_to_all_single(recv_keys, send_keys, output_split_sizes=rc, input_split_sizes=sc, group=self.group)
        masks_here = self.probe_fn(recv_keys)
        back = torch.empty((n, 3), dtype=torch.int64, device=dev)
        dist.all_to_all_single(back, masks_here.contiguous(), output_split_sizes=sc, input_split_sizes=rc,
                               group=self.group)
        out = torch.empty_like(back)
        out[order] = back
        return out


class ShardedIndex:
    """The device path: one Ingest handle per rank, its index holding only this rank's hash range."""

    def __init__(self, ingest, group=None):
        self.h = ingest
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.ex = ShardedExchange(self._probe, group)

    def _probe(self, keys):
        masks = torch.empty((keys.shape[0], 3), dtype=torch.int64, device=keys.device)
        if keys.shape[0]:
            self.h.index_probe_device(keys.data_ptr(), keys.shape[0], masks.data_ptr(),
                                      torch.cuda.current_stream().cuda_stream or None)
        return masks

    def apply(self, instance_id, stored=None, offload=None, removed=None):
        """record_updated_kvcaches for the keys this rank owns (every rank sees the same event stream)."""
        def mine(k):
            if k is None:
                return None
            k = np.asarray(k, dtype=np.uint8).reshape(-1, 16)
            return k[owner_of_numpy(k, self.world) == self.rank]
        self.h.index_apply(instance_id, mine(stored), mine(offload), mine(removed))

    def publish(self):
        self.h.index_publish()

    def match_route(self, d_keys, d_key_start, d_n_blocks, n_req, d_match, d_routing):
        """d_keys uint8 [n_keys,16]; one all-to-all out, local probe, one all-to-all back, then the scan."""
        torch.cuda.current_stream().synchronize()
        masks = self.ex.lookup(d_keys)
        self.h.score_route_device(n_req, masks.data_ptr(), d_key_start.data_ptr(), d_n_blocks.data_ptr(),
                                  d_match.data_ptr() if d_match is not None else None,
                                  d_routing.data_ptr() if d_routing is not None else None,
                                  torch.cuda.current_stream().cuda_stream or None)
        return masks
