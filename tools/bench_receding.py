"""Receding-horizon (closed-loop, perfect tracking) run through the C++ host layer: every MPC cycle re-solves all instances from the state the
previous plan reaches one shooting interval later, warm-started from the previous primal solution (the reference's primalSolution_ path,
SqpSolver.cpp:211-219, Initialization.cpp:35-79).  Reports cold-start vs steady-state throughput and the accepted step sizes.

  python tools/bench_receding.py --batch 256 --cycles 20"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from wb_humanoid_mpc_b200 import abi, host_lib, model_loader  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--cycles", type=int, default=20)
    ap.add_argument("--horizon", type=float, default=3.5)
    ap.add_argument("--gait", default="walk")
    ap.add_argument("--no-spread", action="store_true", help="skip the reference's trajectorySpread of the previous solution")
    args = ap.parse_args()
    model = model_loader.load_packaged_model()
    hm = host_lib.HostModel()
    st = abi.default_settings(model, sqp_iteration=1)
    B, T, dt = args.batch, args.horizon, model["sqp"]["dt"]
    rng = np.random.default_rng(1234)
    solver = host_lib.HostSqpSolver(hm, st, B)
    solver.set_trajectory_spread(not args.no_spread)
    lo, hi = np.array(model["q_lower"]), np.array(model["q_upper"])
    x, cmds = [], []
    for b in range(B):
        x0 = np.array(model["x_init"], float)
        x0[2] = model["reference"]["defaultBaseHeight"]
        x0[0:3] += rng.uniform(-0.02, 0.02, 3)
        x0[3:6] += rng.uniform(-0.05, 0.05, 3)
        x0[6:29] = np.clip(x0[6:29] + rng.uniform(-0.1, 0.1, 23), lo + 0.05, hi - 0.05)
        x0[29:] += rng.uniform(-0.2, 0.2, 29)
        x.append(x0)
        cmds.append([rng.uniform(-0.5, 1.0), rng.uniform(-0.3, 0.3), model["reference"]["defaultBaseHeight"], rng.uniform(-0.5, 0.5)])
        solver.set_gait(b, args.gait, 0.0, args.cycles * dt + 3 * T)
    x = np.array(x)
    rows = []
    t = 0.0
    for c in range(args.cycles):
        for b in range(B):
            solver.set_command(b, t, x[b], cmds[b], T)
        t0 = time.perf_counter()
        solver.run(t, x, t + T)
        wall = time.perf_counter() - t0
        logs = np.array([solver.iterations_log(b)[0] for b in range(B)])
        steps = logs[:, 6]
        rows.append({"cycle": c, "t": round(t, 4), "solves_per_s": B / wall, "ms": 1e3 * wall, "stage_ms": [round(float(v), 3) for v in solver.benchmarks()[:3]],
                     "mean_step": float(steps.mean()), "full_steps": int((steps == 1.0).sum()), "zero_steps": int((steps == 0.0).sum()),
                     "merit_mean": float(logs[:, 3].mean()), "dyn_sse_mean": float(logs[:, 4].mean()), "eq_sse_mean": float(logs[:, 5].mean())})
        # perfect tracking: the next measured state is where the plan is one shooting interval later (node 1 is never an event node here)
        x = np.array([solver.primal_solution(b)["x"][1] for b in range(B)])
        t += dt
    steady = rows[len(rows) // 2:]
    out = {"metric": "SQP solves/sec (G1 whole-body, N=100, batched), receding horizon through b200sqp::host::SqpSolver::run", "unit": "solves/s",
           "cold_start": rows[0]["solves_per_s"], "steady_state": float(np.mean([r["solves_per_s"] for r in steady])),
           "steady_stage_ms": [float(np.mean([r["stage_ms"][k] for r in steady])) for k in range(3)],
           "config": {"batch": B, "cycles": args.cycles, "gait": args.gait, "horizon": T, "trajectory_spread": not args.no_spread}, "cycles": rows}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
