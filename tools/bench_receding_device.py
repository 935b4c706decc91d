"""Device-resident receding horizon: `--cycles` MPC cycles of a batch of whole-body instances with perfect tracking -- every cycle
b200sqp_build_instances(warm = 1, x0 = NULL) shifts the previous solution, interpolates the new measured state from it and rebuilds the references ON THE
DEVICE, then b200sqp_solve runs one SQP iteration (the shipped real-time iteration).  No state or input trajectory crosses PCIe inside the loop; per cycle the
host sends 44 bytes per instance (gait id, gait start, command) and reads nothing.  Prints one JSON line."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from wb_humanoid_mpc_b200 import abi, model_loader  # noqa: E402
from wb_humanoid_mpc_b200.solver import B200SqpSolver  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--cycles", type=int, default=120)
ap.add_argument("--horizon", type=float, default=3.5)
ap.add_argument("--mpc-dt", type=float, default=1.0 / 60.0, help="time between MPC cycles (mpcDesiredFrequency 60 Hz, task.info:115)")
ap.add_argument("--gait", default="walk")
args = ap.parse_args()

import torch  # noqa: E402

model = model_loader.load_packaged_model()
st = abi.default_settings(model, sqp_iteration=1)
insts = bench.build_batch(model, args.batch, 0, args.horizon, [args.gait])
x0 = np.array([i["x0"] for i in insts])
cmd = np.array([i["cmd"] for i in insts])
start = np.zeros(args.batch)
gaits = [args.gait] * args.batch
solver = B200SqpSolver(model, st)
n0 = solver.build_instances(0.0, args.horizon, x0, gaits, start, cmd)
solver.solve()
first = solver.iterations_log()[:, 0]
node_counts, ls_ms = {n0}, []
torch.cuda.synchronize()
t_begin = time.perf_counter()
for c in range(1, args.cycles + 1):
    n = solver.build_instances(c * args.mpc_dt, args.horizon, None, gaits, start, cmd, warm=True)
    node_counts.add(n)
    solver.solve()
    ls_ms.append(solver.benchmarks()[2])
torch.cuda.synchronize()
dt = time.perf_counter() - t_begin
last = solver.iterations_log()[:, 0]
sol = solver.primal_solution()
g0 = np.sqrt(first[:, 2] + first[:, 3])
g1 = np.sqrt(last[:, 2] + last[:, 3])
print(json.dumps({"metric": "SQP solves/sec (G1 whole-body, N=100, batched), receding horizon on the device", "value": args.batch * args.cycles / dt, "unit": "solves/s",
                  "cycles": args.cycles, "batch": args.batch, "ms_per_cycle": 1e3 * dt / args.cycles, "node_counts_seen": sorted(int(v) for v in node_counts),
                  "h2d_bytes_per_cycle": int(args.batch * (4 + 8 + 32)), "d2h_bytes_per_cycle": 12,
                  "constraint_violation_first_cycle_median": float(np.median(g0)), "constraint_violation_last_cycle_median": float(np.median(g1)),
                  "full_steps_last_cycle": int((last[:, 8] == 1.0).sum()), "linesearch_ms_first_last": [float(ls_ms[0]), float(ls_ms[-1])],
                  "status_ok": bool(not sol["status"].any())}))
