"""Galaxy sharding across the GPUs of one box.

Stamps are independent in every hot-path op, so the batch is split into contiguous index ranges, one per rank, with
NO collective on the data path; the only communication is one all-gather of the per-galaxy ellipticities
(8 bytes per galaxy) after the last chunk.  One process per GPU, torch.distributed (NCCL on GPUs, gloo in the CPU
tests of this host logic).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous range [lo, hi) of rank `rank`: ceil(N/R) stamps per rank, the last ranks may be short or empty."""
    per = -(-n_total // world)
    lo = min(n_total, rank * per)
    return lo, min(n_total, lo + per)


def gather_ellipticities(e_local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """[n_r, 2] per-rank moment ellipticities -> [n_total, 2] on every rank, in galaxy-index order.
    Shards are padded to the common ceil(N/R) length so a single fixed-size all_gather suffices."""
    if not (dist.is_available() and dist.is_initialized()):
        assert e_local.shape[0] == n_total
        return e_local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = -(-n_total // world)
    lo, hi = shard_range(n_total, rank, world)
    assert e_local.shape == (hi - lo, 2), (tuple(e_local.shape), lo, hi)
    send = torch.zeros(per, 2, dtype=e_local.dtype, device=e_local.device)
    send[:hi - lo] = e_local
    recv = torch.empty(world * per, 2, dtype=e_local.dtype, device=e_local.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    return recv[:n_total] if world * per == n_total else torch.cat(
        [recv[r * per:r * per + (shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0])] for r in range(world)])


def run_sharded(fn, n_total: int, group=None):
    """Apply fn(lo, hi) -> [hi-lo, 2] ellipticities to this rank's shard and gather everyone's."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_range(n_total, rank, world)
    return gather_ellipticities(fn(lo, hi), n_total, group)
