"""N > 1 path on CPU: world_size-2 gloo processes exercise the sharding rule, the max-over-ranks timing reduction and the global
line-search step selection (the only collectives of the multi-GPU design)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wb_humanoid_mpc_b200 import parallel


def test_shard_bounds_cover_the_batch_exactly():
    for total in (1, 7, 256, 4096, 8192):
        for world in (1, 2, 3, 4, 8):
            seen = np.zeros(total, dtype=int)
            for r in range(world):
                lo, hi = parallel.shard_bounds(total, r, world)
                seen[lo:hi] += 1
                for b in range(lo, hi):
                    assert parallel.owner_of(b, total, world) == r
            assert (seen == 1).all()
    assert parallel.shard_bounds(8192, 3, 8) == (3072, 4096)


def test_alpha_ladder_matches_reference_backtracking():
    lad = parallel.alpha_ladder(0.5, 1e-4)
    assert lad[0] == 1.0 and len(lad) == 14 and lad[-1] >= 1e-4 and lad[-1] * 0.5 < 1e-4


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # (a) timing reduction: max over ranks
        t = parallel.max_over_ranks(1.0 + rank)
        # (b) global step: rank 0's instances accept at alpha index 1, rank 1's only at index 2
        local_batch = 4
        accept = np.array([0, 4, 4, 4], float) if rank == 0 else np.array([0, 1, 4, 4], float)
        merit = np.full(4, 10.0 * (rank + 1))
        vio = np.array([1.0, 0.5, 0.1 * (rank + 1), 0.01])
        idx, acc, mer, v = parallel.select_global_step(accept, merit, vio, local_batch, quorum=1.0)
        idx_q, *_ = parallel.select_global_step(accept, merit, vio, local_batch, quorum=0.6)
        out.put((rank, t, idx, acc.tolist(), mer.tolist(), v.tolist(), idx_q))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_collectives_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=90) for _ in range(world))
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    for rank, t, idx, acc, mer, v, idx_q in res:
        assert t == 2.0                      # max over ranks
        assert idx == 2                      # largest alpha accepted by ALL 8 instances
        assert acc == [0.0, 5.0, 8.0, 8.0]
        assert mer == [30.0] * 4
        assert v == [1.0, 0.5, 0.2, 0.01]
        assert idx_q == 1                    # 60 % quorum is already met at index 1 (5 of 8)
