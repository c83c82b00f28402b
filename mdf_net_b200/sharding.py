"""Multi-GPU plumbing of the path: shard independent reference views over ranks, gather the final maps.

The plane-sweep path has no data-path collective: one unit of work is one (scan, reference view)
sample -- one `Dataset.__getitem__` of the reference (load/dtueval.py:23-61) -- and samples share no
state in eval (eval.py:23-24).  Rank r of R takes the units i with i % R == r (round robin keeps the
ranks balanced when scans have 49 views); the only exchange is the optional gather of the final
depth / confidence maps (2 x H0 x W0 floats per view) to rank 0, which uses `torch.distributed`
(NCCL over NVLink on GPUs, gloo on CPU for the tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_units(num_units: int, rank: int, world_size: int) -> List[int]:
    """Indices of the (scan, reference view) units rank `rank` processes."""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} / world size {world_size}")
    return list(range(rank, max(num_units, 0), world_size))


def units_of_scans(views_per_scan: Sequence[int]) -> List[Tuple[int, int]]:
    """Flat unit list [(scan, ref_view), ...] in the order the reference's eval loader walks it
    (load/dtueval.py:51-61: scans outer, pair.txt reference views inner)."""
    return [(s, v) for s, n in enumerate(views_per_scan) for v in range(n)]


def gather_maps(local: torch.Tensor, num_units: int, group: Optional[dist.ProcessGroup] = None,
                dst: int = 0) -> Optional[torch.Tensor]:
    """Gather per-unit maps to rank `dst`.

    local: (n_local, ...) maps of this rank's units, in the order of `shard_units`.
    Returns on `dst` a (num_units, ...) tensor in global unit order, None elsewhere.
    Ranks may own different numbers of units (49 views over 8 GPUs): shards are padded to the largest.
    """
    if not dist.is_available() or not dist.is_initialized():
        if local.shape[0] != num_units:
            raise ValueError("single process: local must hold every unit")
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per_rank = [len(range(r, num_units, world)) for r in range(world)]
    width = max(per_rank) if per_rank else 0
    if local.shape[0] != per_rank[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} units, expected {per_rank[rank]}")
    padded = local.new_zeros((width,) + tuple(local.shape[1:]))
    padded[: local.shape[0]] = local
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    out = local.new_empty((num_units,) + tuple(local.shape[1:]))
    for r in range(world):
        out[r::world] = bufs[r][: per_rank[r]]
    return out
