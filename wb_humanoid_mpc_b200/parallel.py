"""Multi-GPU plumbing: batch-index sharding of independent MPC instances, one process per GPU.

The data path needs no collective: instance b of a global batch of B goes to rank floor(b*G/B) and every rank runs the whole SQP
iteration on its shard.  Collectives are used only for (a) the max-over-ranks of timings and (b) the optional "global step" mode,
where all ranks agree on one line-search step size from per-candidate acceptance statistics (one all-reduce of <= 14 x 3 doubles).
torch.distributed is plumbing only (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(total: int, rank: int, world: int) -> tuple[int, int]:
    """contiguous slice [lo, hi) of the global batch owned by `rank`: instance b -> rank floor(b*world/total)"""
    lo = -(-rank * total // world)          # ceil(rank*total/world)
    hi = -(-(rank + 1) * total // world)
    return lo, hi


def owner_of(b: int, total: int, world: int) -> int:
    return b * world // total


def alpha_ladder(alpha_decay: float = 0.5, alpha_min: float = 1e-4) -> np.ndarray:
    """candidate step sizes 1, decay, decay^2, ... >= alpha_min (SqpSolver::takeStep back-tracking, SqpSolver.cpp:517-565)"""
    out, a = [], 1.0
    while a >= alpha_min:
        out.append(a)
        a *= alpha_decay
    return np.array(out)


def max_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def select_global_step(accept_count, merit_sum, max_violation, local_batch: int, quorum: float = 1.0, device=None):
    """Global line-search step: every rank evaluated the fixed ladder for its instances and passes, per candidate alpha,
    [#instances that would accept, sum of merits, max constraint violation].  One all-reduce (SUM, SUM, MAX packed as two calls on
    one small tensor each) and every rank picks the same alpha* = the largest candidate accepted by >= quorum of all instances
    (0.0 when none is).  Returns (alpha_index or -1, global_accept_count, global_merit_sum, global_max_violation)."""
    import torch
    import torch.distributed as dist

    acc = torch.as_tensor(np.asarray(accept_count, dtype=np.float64), device=device)
    mer = torch.as_tensor(np.asarray(merit_sum, dtype=np.float64), device=device)
    vio = torch.as_tensor(np.asarray(max_violation, dtype=np.float64), device=device)
    tot = torch.tensor([float(local_batch)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        packed = torch.cat([acc, mer, tot])
        dist.all_reduce(packed, op=dist.ReduceOp.SUM)
        acc, mer, tot = packed[: len(acc)], packed[len(acc): 2 * len(acc)], packed[-1:]
        dist.all_reduce(vio, op=dist.ReduceOp.MAX)
    need = quorum * float(tot.item())
    idx = -1
    for i, a in enumerate(acc.tolist()):
        if a >= need - 1e-9:
            idx = i
            break
    return idx, acc.cpu().numpy(), mer.cpu().numpy(), vio.cpu().numpy()
