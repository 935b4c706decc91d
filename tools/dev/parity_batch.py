"""Development aid: per-instance GPU-vs-oracle differences over the benchmark batch (which instances differ most, and in which block)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import bench  # noqa: E402
import oracle_lib as orc  # noqa: E402
from test_gpu_wb import oracle_solve, rel  # noqa: E402
from wb_humanoid_mpc_b200 import abi, model_loader  # noqa: E402
from wb_humanoid_mpc_b200.solver import B200SqpSolver, stack_instances  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
model = model_loader.load_packaged_model()
st = abi.default_settings(model, sqp_iteration=1)
insts = bench.build_batch(model, n, 0, 3.5, ["walk"])
solver = B200SqpSolver(model, st, capture_raw_blocks=True)
sol = solver.run(stack_instances(insts))
raw = solver.raw_stage_blocks()
import os
dt, xo, uo = bench.oracle_batch_solve(model, insts, st, os.cpu_count())
ex = np.abs(sol["x"] - xo).reshape(n, -1).max(1)
eu = np.abs(sol["u"] - uo).reshape(n, -1).max(1)
order = np.argsort(-ex)
print("oracle", dt, "s; worst instances (idx, |dx|, |du|, alpha):")
for i in order[:8]:
    print(" ", i, f"{ex[i]:.2e} {eu[i]:.2e}", sol["log"][i, 0, 8])
print("median", np.median(ex), "p90", np.quantile(ex, 0.9))
w = int(order[0])
inst = insts[w]
ref = oracle_solve(model, inst, st, keep_raw=True)
nn = len(inst["t_nodes"])
print("worst instance", w, "single-oracle vs batch-oracle x diff", np.abs(ref["x"] - xo[w]).max())
for k in range(nn - 1):
    g = orc.unpack_raw_blocks(raw[w, k], 58, 35)
    o = ref["raw"][k]
    bad = {key: rel(g[key], o[key]) for key in ["A", "B", "b", "Q", "S", "R", "q", "r", "C", "D", "e"] if rel(g[key], o[key]) > 1e-11}
    if bad or g["nc"] != o["nc"]:
        print("  stage", k, "nc", g["nc"], o["nc"], bad)
alpha = sol["log"][w, 0, 8]
dx = (sol["x"][w] - inst["x_init"]) / alpha
du = (sol["u"][w] - inst["u_init"]) / alpha
per = [np.max(np.abs(dx[k] - ref["dx"][k])) for k in range(nn)]
peru = [np.max(np.abs(du[k] - ref["du"][k])) for k in range(nn - 1)]
print("  dx err per node (every 8th):", [f"{per[k]:.1e}" for k in range(0, nn, 8)])
print("  du err per node (every 8th):", [f"{peru[k]:.1e}" for k in range(0, nn - 1, 8)], "argmax", int(np.argmax(peru)), f"{max(peru):.2e}")
print("  log gpu", sol["log"][w, 0, :13])
print("  log orc", ref["log"][0][:13])
# conditioning of the projected input Hessian along the horizon (oracle blocks): cond(D) of the constraint Jacobian
cd = []
for k in range(nn - 1):
    o = ref["raw"][k]
    if o["nc"]:
        s = np.linalg.svd(o["D"], compute_uv=False)
        cd.append(s[0] / s[-1])
print("  cond(D) max", f"{max(cd):.2e}", "median", f"{np.median(cd):.2e}")
