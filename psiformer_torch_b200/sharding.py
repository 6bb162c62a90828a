"""Walker sharding over the GPUs of one box.

Walkers are independent Markov chains and independent E_L evaluations, so the data path needs no
collective: global walker ids [0, W) are split contiguously, one rank per GPU.  The only exchange is
the reduction that ``Trainer.train`` performs at train.py:138 (mean local energy) and the gradient
mean after ``loss.backward()``; both are ``torch.distributed`` all-reduces (NCCL over NVLink on the
GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, Tuple

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class WalkerShard:
    rank: int
    world_size: int
    total_walkers: int

    @property
    def bounds(self) -> Tuple[int, int]:
        """[lo, hi) global walker ids of this rank (sizes differ by at most one)."""
        q, r = divmod(self.total_walkers, self.world_size)
        lo = self.rank * q + min(self.rank, r)
        return lo, lo + q + (1 if self.rank < r else 0)

    @property
    def walker_id0(self) -> int:
        return self.bounds[0]

    @property
    def local_walkers(self) -> int:
        lo, hi = self.bounds
        return hi - lo


def current_shard(total_walkers: int) -> WalkerShard:
    if dist.is_available() and dist.is_initialized():
        return WalkerShard(dist.get_rank(), dist.get_world_size(), total_walkers)
    return WalkerShard(0, 1, total_walkers)


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def allreduce_energy_stats(accum: torch.Tensor) -> torch.Tensor:
    """Sum {sum E_L, sum E_L^2, n} (fp64[3], as accumulated on the device by psif_local_energy)
    over ranks.  Returns the same tensor."""
    assert accum.dtype == torch.float64 and accum.numel() == 3
    if is_distributed():
        dist.all_reduce(accum, op=dist.ReduceOp.SUM)
    return accum


def agree_max(value: float, device: torch.device | None = None) -> float:
    """Maximum of a host number over the ranks.  Anything that decides HOW MANY collectives a rank will issue (a loop
    bound derived from a local timing, a "keep going" flag) has to go through this first: ranks that disagree by one
    iteration pair an all-reduce with the next, differently sized, collective and the job hangs."""
    if not is_distributed():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or torch.device("cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def energy_mean_and_variance(accum: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    n = accum[2].clamp_min(1.0)
    mean = accum[0] / n
    return mean, (accum[1] / n - mean * mean).clamp_min(0.0)


def allreduce_mean_gradients(params: Iterable[torch.nn.Parameter]) -> None:
    """Average parameter gradients over ranks with one flat all-reduce (3.2 M fp32 = 12.8 MB)."""
    if not is_distributed():
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= dist.get_world_size()
    o = 0
    for g in grads:
        g.copy_(flat[o:o + g.numel()].view_as(g))
        o += g.numel()


def broadcast_parameters(params: Iterable[torch.nn.Parameter], src: int = 0) -> None:
    """Make every replica start from rank ``src``'s parameters (the reference seeds nothing: train.py:247-265)."""
    if not is_distributed():
        return
    with torch.no_grad():
        for p in params:
            dist.broadcast(p.data, src=src)
