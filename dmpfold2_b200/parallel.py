"""Multi-GPU plumbing: independent alignments are sharded one-per-GPU, one process per GPU, with NO data-path
collective.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used only to
agree on the shard, to gather the small per-target results and to take the max-over-ranks of a device time.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def targets_for_rank(num_targets: int, rank: int, world_size: int) -> List[int]:
    """Round-robin shard: target t belongs to rank t % world_size (equal work per rank up to one target)."""
    if not (0 <= rank < world_size):
        raise ValueError(f'rank {rank} outside world of {world_size}')
    return list(range(rank, num_targets, world_size))


def max_over_ranks(value: float, device: Optional[torch.device] = None) -> float:
    """Max of a per-rank scalar (e.g. a CUDA-event duration) over all ranks."""
    rank, ws = world()
    if ws == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or 'cpu')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def fold_many(targets: Sequence, fold_fn: Callable[[object], object], gather: bool = True) -> List[object]:
    """Fold a list of targets across the ranks of the current process group.

    Every rank calls this with the SAME `targets` list; rank r runs `fold_fn` on targets r, r+W, r+2W, ... (no
    communication while folding).  With gather=True every rank returns the full result list in target order
    (results are small: (L,5,3) coordinates and (L,) confidences); otherwise only the local results, with None
    for targets folded elsewhere.
    """
    rank, ws = world()
    mine = targets_for_rank(len(targets), rank, ws)
    local = {t: fold_fn(targets[t]) for t in mine}
    if ws == 1 or not gather:
        return [local.get(t) for t in range(len(targets))]
    parts: List[Optional[dict]] = [None] * ws
    dist.all_gather_object(parts, local)
    merged = {}
    for p in parts:
        merged.update(p)
    return [merged[t] for t in range(len(targets))]
