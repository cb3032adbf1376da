"""Multi-GPU plumbing: shard target views over ranks, gather rendered frames.

The path is embarrassingly parallel over target views — exactly how the reference
parallelises (one process per GPU, `DistributedSampler(shuffle=False)`, 1 view per GPU per
step: /root/reference/pgdvs/engines/trainer_pgdvs.py:290-306, scripts/benchmark.sh:316).
There is no exchange inside the path, so no data-path collective; NCCL (over NVLink, P2P left
enabled — the reference sets NCCL_P2P_DISABLE=1, scripts/benchmark.sh:36) is used only to
gather finished frames on one rank, e.g. for the video writer
(/root/reference/pgdvs/engines/visualizer_pgdvs.py:160-177 globs per-rank PNGs from disk instead).
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


def shard_views(n_views: int, rank: int, world_size: int, pad: bool = False) -> List[int]:
    """View indices of `rank`: v -> rank (v mod world_size), like DistributedSampler(shuffle=False).
    With pad=True the tail is padded by wrapping around (DistributedSampler's drop_last=False
    behaviour) so that every rank gets the same count."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    idx = list(range(n_views))
    if pad and n_views % world_size:
        total = ((n_views + world_size - 1) // world_size) * world_size
        idx += idx[: total - n_views]
    return idx[rank::world_size]


def gather_frames(local: torch.Tensor, n_views: int, dst: int = 0, group=None,
                  out_list: Optional[List[torch.Tensor]] = None) -> Optional[torch.Tensor]:
    """Gather per-rank frames [V_local, ...] (sharded with shard_views(..., pad=True)) on `dst`
    and restore the global view order.  Returns [n_views, ...] on dst, None elsewhere.
    Works with NCCL (CUDA tensors) and gloo (CPU tensors; used by the CPU tests)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return local[:n_views]
    if rank == dst:
        bufs = out_list if out_list is not None else [torch.empty_like(local) for _ in range(world)]
        dist.gather(local, bufs, dst=dst, group=group)
        # rank r holds views r, r + world, ...  ->  interleave
        stacked = torch.stack(bufs, dim=1)  # [V_local, world, ...]
        return stacked.reshape((-1,) + tuple(local.shape[1:]))[:n_views]
    dist.gather(local, None, dst=dst, group=group)
    return None
