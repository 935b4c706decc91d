"""Development tool: time build variants of libb200sqp.so (tools/_exp/lib_<name>.so, built by hand with -DB200SQP_*_THREADS=... etc.) on the
bench workload (batch 256, walk, 115 nodes, cold start) and check that every variant's primal solution equals the shipped library's.
Usage: python tools/dev/variant_sweep.py [name ...]        (worker: python tools/dev/variant_sweep.py --worker <name>)"""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))


def worker(name, steps=8, warmup=3):
    import numpy as np
    import torch

    import bench
    from wb_humanoid_mpc_b200 import abi, model_loader
    from wb_humanoid_mpc_b200.solver import B200SqpSolver, stack_instances

    model = model_loader.load_packaged_model()
    settings = abi.default_settings(model, sqp_iteration=1, global_step=0)
    batch = stack_instances(bench.build_batch(model, 256, 0, 3.5, ["walk"]))
    solver = B200SqpSolver(model, settings, device=0)
    solver.upload({k: (v if v.dtype == np.uint8 else v.astype(np.float64)) for k, v in batch.items()})
    for _ in range(warmup):
        solver.reset()
        solver.solve()
    acc = np.zeros(4)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        solver.reset()
        solver.solve()
        acc += np.array(solver.benchmarks())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    sol = solver.primal_solution()
    np.save(f"/tmp/x_{name}.npy", sol["x"])
    print(json.dumps(dict(name=name, solves_per_s=256e3 / ms, ms_per_step=ms, lq=acc[0] / steps, lq_proj=acc[3] / steps, qp=acc[1] / steps,
                          linesearch=acc[2] / steps)))


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--worker":
        return worker(sys.argv[2])
    import numpy as np

    names = sys.argv[1:] or sorted(p.stem[4:] for p in (ROOT / "tools" / "_exp").glob("lib_*.so"))
    ref = None
    for name in ["shipped"] + names:
        env = dict(os.environ)
        if name != "shipped":
            env["B200SQP_LIB"] = str(ROOT / "tools" / "_exp" / f"lib_{name}.so")
        try:
            res = subprocess.run([sys.executable, __file__, "--worker", name], env=env, capture_output=True, text=True, timeout=300)
        except subprocess.TimeoutExpired:
            print(json.dumps(dict(name=name, error="timeout")))
            continue
        line = [l for l in res.stdout.splitlines() if l.startswith("{")]
        if not line:
            print(json.dumps(dict(name=name, error=(res.stderr or res.stdout)[-600:])))
            continue
        r = json.loads(line[-1])
        x = np.load(f"/tmp/x_{name}.npy")
        if name == "shipped":
            ref = x
        elif ref is not None:
            r["max_abs_diff_x_vs_shipped"] = float(np.abs(x - ref).max())
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
