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


def step_seed(it: int, rank: int = 0, world: int = 1) -> int:
    """Dropout seed of meta-step ``it`` on ``rank``.  The engine derives task t's Philox keys as seed * 128 + 2 * slot +
    pass with the rank-LOCAL slot, so the seed itself has to differ between ranks (else slot 0 of every rank -- tasks
    0, 1, 2, ... -- would draw identical masks) and between runs (torch.initial_seed(), i.e. args.seed through
    torch.manual_seed as in meta_transfer_train.py:109-112).  48 bits: seed * 128 stays inside a u64."""
    mix = (torch.initial_seed() * 0x9E3779B97F4A7C15) & ((1 << 64) - 1)
    return ((mix >> 16) ^ (it * world + rank)) & ((1 << 48) - 1)


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


class MetaExchange:
    """The end of a sharded meta-step: exchange of copy_grad + the outer optimizer step (transient_trainer.py:248-255).

    mode "allreduce": all-reduce(SUM) of the flat arena, then the identical full Adam step on every rank
                      (``mtl_meta_finish``); .grad holds the full outer gradient as in the reference.  With
                      ``overlap`` (default) the exchange is split where the data becomes final: region A -- every
                      parameter but the VGG front-end, 98 % of the bytes -- is all-reduced on a side stream as soon as
                      the last task's validation pass reaches its VGG backward (``mtl_stream_wait_region_a``), i.e.
                      under the ~0.4 ms of convolution backward that is still running; only the 1 MB VGG tail is
                      exchanged after the step.  Same sums, same order on every rank.
    mode "sharded"  : reduce-scatter(SUM) of the arena -> Adam on this rank's 1/world slice of theta / m / v ->
                      all-gather of the theta slices.  Same bytes on the wire as the all-reduce (which is a
                      reduce-scatter + all-gather inside NCCL), 1/world of the optimizer's HBM traffic, and the
                      all-gather moves finished weights instead of gradients.  m / v are only maintained for the
                      rank's slice (``gather_moments`` rebuilds the full arenas for a checkpoint).  Needs
                      n_floats % world == 0 and no clipping (the clip coefficient needs the norm of the full
                      gradient); otherwise this step falls back to "allreduce".
    Without a process group both reduce to ``Session.meta_finish``."""

    def __init__(self, session, dist=None, mode: str = "allreduce", overlap: bool = True):
        if mode not in ("allreduce", "sharded"):
            raise ValueError(f"unknown exchange mode {mode!r}")
        self.s, self.dist, self.mode, self.overlap = session, dist, mode, overlap
        self.world = dist.get_world_size() if dist is not None else 1
        self.rank = dist.get_rank() if dist is not None else 0
        self._shard = None
        self._theta_shard = None
        self._sharded_last = False
        self._side = None
        self.ran_tasks = True      # set False by a caller whose rank had no task this step (nothing signalled region A)

    def _slice(self, arena):
        n = arena.numel() // self.world
        return arena[self.rank * n:(self.rank + 1) * n]

    def finish(self, theta, grad, copy_grad, adam_m, adam_v, adam_state, meta_lr, clip=False, max_norm=400.0,
               betas=(0.9, 0.999), eps=1e-8):
        s, d = self.s, self.dist
        sharded = (d is not None and self.mode == "sharded" and not clip and copy_grad.numel() % self.world == 0)
        self._sharded_last = sharded
        if not sharded:
            if d is not None and self.overlap and copy_grad.is_cuda:
                n_a = s.region_a_floats()
                if self._side is None:
                    self._side = torch.cuda.Stream(device=copy_grad.device)
                cur = torch.cuda.current_stream(copy_grad.device)
                if self.ran_tasks:
                    s.wait_region_a(self._side)            # not the whole step: only region A of this step's copy_grad
                else:
                    self._side.wait_stream(cur)
                with torch.cuda.stream(self._side):
                    work_a = d.all_reduce(copy_grad[:n_a], op=d.ReduceOp.SUM, async_op=True)
                d.all_reduce(copy_grad[n_a:], op=d.ReduceOp.SUM)       # the VGG tail, after the step
                work_a.wait()
                cur.wait_stream(self._side)
            else:
                exchange_copy_grad(copy_grad, d)
            s.meta_finish(theta, grad, copy_grad, adam_m, adam_v, adam_state, meta_lr, clip=clip, max_norm=max_norm,
                          betas=betas, eps=eps)
            return
        n = copy_grad.numel() // self.world
        if self._shard is None or self._shard.numel() != n:
            self._shard = torch.empty(n, dtype=torch.float32, device=copy_grad.device)
            self._theta_shard = torch.empty(n, dtype=torch.float32, device=copy_grad.device)
        d.reduce_scatter_tensor(self._shard, copy_grad, op=d.ReduceOp.SUM)
        self._theta_shard.copy_(self._slice(theta))
        s.adam(self._theta_shard, self._shard, self._slice(adam_m), self._slice(adam_v), adam_state, meta_lr,
               betas[0], betas[1], eps)
        d.all_gather_into_tensor(theta, self._theta_shard)

    def last_copy_grad(self, copy_grad):
        """The summed outer gradient of the last finish() as one full arena (a collective in sharded mode: testing and
        logging only)."""
        if not self._sharded_last:
            return copy_grad
        full = torch.empty_like(copy_grad)
        self.dist.all_gather_into_tensor(full, self._shard)
        return full

    def gather_moments(self, adam_m, adam_v):
        """Sharded mode: make the full m / v arenas valid on every rank (before saving a checkpoint)."""
        if self.dist is None or self.mode != "sharded":
            return
        for a in (adam_m, adam_v):
            sl = self._slice(a).clone()
            self.dist.all_gather_into_tensor(a, sl)
