"""BASELINE configs[4]: mixed contact-schedule sweep (stance / walk / slow_walk with random gait phase offsets) through the C++ host layer
(b200sqp::host::SqpSolver groups the instances by node count, one library handle per group).

  python tools/bench_mixed.py --batch 1024 --steps 5        # one GPU; under torchrun every rank solves its own shard (batch per GPU)
Prints one JSON line (same keys as bench.py where they apply); the timed region is SqpSolver::run = host instance building + upload + solve +
download, i.e. an end-to-end number."""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from wb_humanoid_mpc_b200 import abi, host_lib, model_loader  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--horizon", type=float, default=3.5)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    import torch

    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    model = model_loader.load_packaged_model()
    hm = host_lib.HostModel()
    st = abi.default_settings(model, sqp_iteration=1)
    rng = np.random.default_rng(1234 + 7919 * rank)
    B = args.batch
    gaits = ["stance", "walk", "slow_walk"]
    solver = host_lib.HostSqpSolver(hm, st, B, device=local)
    lo, hi = np.array(model["q_lower"]), np.array(model["q_upper"])
    x0s = []
    for b in range(B):
        x0 = np.array(model["x_init"], float)
        x0[2] = model["reference"]["defaultBaseHeight"]
        x0[0:3] += rng.uniform(-0.02, 0.02, 3)
        x0[3:6] += rng.uniform(-0.05, 0.05, 3)
        x0[6:29] = np.clip(x0[6:29] + rng.uniform(-0.1, 0.1, 23), lo + 0.05, hi - 0.05)
        x0[29:] += rng.uniform(-0.2, 0.2, 29)
        g = gaits[b % 3]
        period = model["gaits"][g]["switchingTimes"][-1] if g != "stance" else 1.0
        solver.set_gait(b, g, -rng.uniform(0.0, period), 3 * args.horizon)      # gait phase offset ~ U[0, period)
        solver.set_command(b, 0.0, x0, [rng.uniform(-0.5, 1.0), rng.uniform(-0.3, 0.3), model["reference"]["defaultBaseHeight"], rng.uniform(-0.5, 0.5)],
                           args.horizon)
        x0s.append(x0)
    x0s = np.array(x0s)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    for _ in range(args.warmup):
        solver.reset()
        solver.run(0.0, x0s, args.horizon)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        solver.reset()
        solver.run(0.0, x0s, args.horizon)
    dt = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    groups = {}
    for b in range(B):
        n = len(solver.primal_solution(b)["t"])
        groups[n] = groups.get(n, 0) + 1
    ms = solver.benchmarks()
    if rank == 0:
        print(json.dumps({"metric": "SQP solves/sec (G1 whole-body, N=100, batched)", "value": B * world * args.steps / dt, "unit": "solves/s", "n_gpus": world,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
                          "dtype": "f64", "data": "synthetic",
                          "config": {"workload": f"G1 whole-body MPC, mixed contact-schedule sweep (stance/walk/slow_walk, random gait phase), batch={B}/GPU, "
                                                 "through b200sqp::host::SqpSolver::run (end to end)",
                                     "node_count_groups": {str(k): v for k, v in sorted(groups.items())}},
                          "stage_ms_last_run": {"lq": ms[0], "qp": ms[1], "linesearch": ms[2]}}))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
