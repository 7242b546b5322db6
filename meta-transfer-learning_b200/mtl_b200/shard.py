"""Task sharding and the one exchange step of the multi-GPU meta-step (host logic, device-agnostic).

SURVEY 8e: given theta0, task i's work (trainer/asr/transient_trainer.py:178-237) reads only theta0 and writes
only its contribution to copy_grad, so tasks shard over ranks with theta / Adam state replicated and ONE
all-reduce(SUM) of the flat copy_grad arena per meta-step; every rank then applies the identical Adam step
(replicas stay bit-identical because the all-reduce result is).  Works on any backend (NCCL on the GPU box,
gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


def dist_env():
    """(dist module or None, rank, world) for the current process."""
    d = torch.distributed
    if d.is_available() and d.is_initialized():
        return d, d.get_rank(), d.get_world_size()
    return None, 0, 1


def task_shard(n_tasks: int, rank: int, world: int) -> List[int]:
    """Tasks (manifest ids) owned by ``rank``: i with i % world == rank, in the reference's task order."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n_tasks, world))


def exchange_copy_grad(copy_grad: torch.Tensor, dist=None) -> torch.Tensor:
    """The exchange step: in-place SUM of the flat copy_grad arena over all ranks (no-op without a group).
    The arena must be one contiguous fp32 tensor so this is a single collective."""
    if copy_grad.dim() != 1 or not copy_grad.is_contiguous() or copy_grad.dtype != torch.float32:
        raise ValueError("copy_grad must be one flat contiguous fp32 arena")
    if dist is not None:
        dist.all_reduce(copy_grad, op=dist.ReduceOp.SUM)
    return copy_grad


def reduce_stats(values: Sequence[float], device, dist=None) -> Tuple[float, ...]:
    """Sum of per-rank logging scalars (loss sum, CER numerator / denominator) over the ranks."""
    if dist is None:
        return tuple(float(v) for v in values)
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t)
    return tuple(float(v) for v in t)
