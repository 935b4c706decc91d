"""Development aid: error budget of one benchmark-size instance (GPU vs oracle), per quantity and per stage.  Run on a GPU box."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib as orc  # noqa: E402
from test_gpu_wb import make_instances, oracle_solve, rel  # noqa: E402
from wb_humanoid_mpc_b200 import abi, model_loader  # noqa: E402
from wb_humanoid_mpc_b200.solver import B200SqpSolver  # noqa: E402

model = model_loader.load_packaged_model()
for gait, cmd, hor in [("walk", [0.5, 0.0, 0.7925, 0.0], 3.5), ("stance", None, 3.5), ("walk", [0.5, 0.0, 0.7925, 0.0], 1.1)]:
    inst = make_instances(model, np.random.default_rng(31), [(gait, hor, cmd)])[0]
    n = len(inst["t_nodes"])
    st = abi.default_settings(model, sqp_iteration=1, use_feedback_policy=1)
    solver = B200SqpSolver(model, st, capture_raw_blocks=True)
    sol = solver.run([inst])
    raw = solver.raw_stage_blocks()
    ref = oracle_solve(model, inst, st, keep_raw=True)
    worst = {}
    for k in range(n - 1):
        g = orc.unpack_raw_blocks(raw[0, k], 58, 35)
        o = ref["raw"][k]
        for key in ["A", "B", "b", "Q", "S", "R", "q", "r", "C", "D", "e"]:
            worst[key] = max(worst.get(key, 0.0), rel(g[key], o[key]))
    g, o = sol["log"][0, 0], ref["log"][0]
    alpha = g[8]
    dx = (sol["x"][0] - inst["x_init"]) / alpha
    du = (sol["u"][0] - inst["u_init"]) / alpha
    print(gait, hor, "nodes", n, "alpha", alpha, o[8])
    print("  blocks", {k: f"{v:.1e}" for k, v in worst.items()})
    print("  log rel", [f"{abs(g[j]-o[j])/max(1.0,abs(o[j])):.1e}" for j in range(13)])
    print(f"  dx {rel(dx, ref['dx']):.2e} du {rel(du, ref['du']):.2e} x {rel(sol['x'][0], ref['x']):.2e} u {rel(sol['u'][0], ref['u']):.2e} "
          f"K {rel(sol['K'][0], ref['K']):.2e} max|x-xo| {np.max(np.abs(sol['x'][0]-ref['x'])):.2e}")
    per = [np.max(np.abs(dx[k] - ref["dx"][k])) for k in range(n)]
    print("  per-node |dx err| first/mid/last", [f"{per[i]:.1e}" for i in (0, 1, n // 4, n // 2, 3 * n // 4, n - 1)], "argmax", int(np.argmax(per)))
    peru = [np.max(np.abs(du[k] - ref["du"][k])) for k in range(n - 1)]
    print("  per-node |du err| argmax", int(np.argmax(peru)), f"{max(peru):.1e}", "max|du|", f"{np.max(np.abs(ref['du'])):.2e}")
