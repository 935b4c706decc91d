"""BASELINE configs[1] (and [0]): G1 full centroidal dynamics, N = 100 (dt 0.02 s, horizon 2.0 s), batch 1 on one B200 -- a correctness
configuration; this script reports its latency next to the CPU oracle (one thread per instance) and re-checks the parity of the very solve
it timed.  `--batch` > 1 gives the throughput of the (correctness-first) centroidal kernels.  One JSON line, same keys as bench.py where
they apply.  The timed region of `value` is b200sqp_reset + b200sqp_solve on device-resident instances (CUDA-event stage times summed by
the library); `e2e` adds upload and download through the C ABI."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--horizon", type=float, default=2.0)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--sqp-iteration", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    import torch

    from wb_humanoid_mpc_b200 import abi, model_loader, references
    from wb_humanoid_mpc_b200.solver import B200SqpSolver, stack_instances

    if not torch.cuda.is_available():
        raise SystemExit("bench_centroidal: no CUDA device; there is no CPU fallback")
    model = model_loader.load_packaged_model("g1_centroidal")
    rng = np.random.default_rng(1234)
    insts = []
    for _ in range(args.batch):
        x0 = np.array(model["x_init"], float)
        x0[6:8] += rng.uniform(-0.02, 0.02, 2)
        x0[9:12] += rng.uniform(-0.05, 0.05, 3)
        x0[12:] += rng.uniform(-0.1, 0.1, model["nj"])
        cmd = [rng.uniform(-0.5, 1.0), rng.uniform(-0.3, 0.3), model["reference"]["defaultBaseHeight"], rng.uniform(-0.5, 0.5)]
        insts.append(references.build_instance(model, x0, gait="walk", cmd=cmd, horizon=args.horizon))
    n_nodes = len(insts[0]["t_nodes"])
    st = abi.default_settings(model, sqp_iteration=args.sqp_iteration)
    solver = B200SqpSolver(model, st)
    batch = stack_instances(insts)
    solver.upload(batch)
    ms_dev, stage = [], np.zeros(4)
    for it in range(args.warmup + args.steps):
        solver.reset()
        torch.cuda.synchronize()
        solver.solve()
        torch.cuda.synchronize()
        if it >= args.warmup:
            b = solver.benchmarks()
            ms_dev.append(b[0] + b[1] + b[2])
            stage += np.array(b)
    stage /= args.steps
    sol = solver.primal_solution()
    e2e_ms = []
    for it in range(args.warmup + args.steps):
        t = time.perf_counter()
        solver.upload(batch)
        solver.solve()
        sol = solver.primal_solution()
        if it >= args.warmup:
            e2e_ms.append((time.perf_counter() - t) * 1e3)
    ms = float(np.mean(ms_dev))
    out = {"metric": "SQP solves/sec (G1 centroidal, N=%d, batched)" % round(args.horizon / model["sqp"]["dt"]), "value": args.batch / ms * 1e3,
           "unit": "solves/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": "G1 centroidal MPC (nx=35, nu=35, FullCentroidalDynamics), dt=0.02 s, horizon %.1f s, gait=walk, batch=%d, "
                                  "sqpIteration=%d, cold start" % (args.horizon, args.batch, args.sqp_iteration), "n_nodes": n_nodes},
           "stage_ms": {"lq": float(stage[0]), "lq_projection_share": float(stage[3]), "qp": float(stage[1]), "linesearch": float(stage[2])},
           "e2e": {"value": args.batch / float(np.mean(e2e_ms)) * 1e3, "unit": "solves/s", "ms_per_step": float(np.mean(e2e_ms))},
           "gpu_launches": solver.launch_count()}
    if not args.no_cpu_baseline:
        import oracle_lib as orc

        o = orc.CenOracle(model)
        inst = insts[0]
        o.set_nodes(inst["contact_flags"], inst["swing_ref"], inst["impact_factor"], inst["arm_phase"], inst["x_ref"])
        t = time.perf_counter()
        ref = o.sqp(inst["t_nodes"], inst["node_event"], inst["x0"], inst["x_init"], inst["u_init"], st)
        dt_cpu = time.perf_counter() - t
        out["cpu_baseline"] = {"value": 1.0 / dt_cpu, "unit": "solves/s", "cores": 1, "kind": "port",
                               "sample": "instance 0 of the batch, one thread (dense forward-mode duals, 70 directions), %.1f s" % dt_cpu,
                               "max_abs_diff_x_vs_gpu": float(np.abs(ref["x"] - sol["x"][0]).max()),
                               "max_abs_diff_u_vs_gpu": float(np.abs(ref["u"] - sol["u"][0]).max())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
