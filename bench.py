#!/usr/bin/env python
"""Benchmark of the hot path: SQP solves/s for the Unitree G1 whole-body OCP (one solve = one SqpSolver::runImpl with sqpIteration = 1).

    python bench.py --gpus N --steps K --warmup W                 # this repo's CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference algorithm on the host cores (fast CPU restatement)

One "step" = one batched solve of `--batch` independent MPC instances per GPU (default 256 = BASELINE.json configs[2]).
Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for the definitions of every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from wb_humanoid_mpc_b200 import abi, model_loader, references  # noqa: E402

METRIC = "SQP solves/sec (G1 whole-body, N=100, batched)"
SEED = 1234


def build_batch(model, batch, rank, horizon, gaits, random_phase=False):
    """Instance distribution of SURVEY.md §8d (seed 1234): perturbed initial states, velocity commands, cold start.
    random_phase: gait phase offset ~ U[0, period) per instance (BASELINE configs[4], the mixed contact-schedule sweep)."""
    rng = np.random.default_rng(SEED + 7919 * rank)
    nj = model["nj"]
    lo, hi = np.array(model["q_lower"]), np.array(model["q_upper"])
    insts = []
    for i in range(batch):
        x0 = np.array(model["x_init"], float)
        x0[2] = model["reference"]["defaultBaseHeight"]
        x0[0:3] += rng.uniform(-0.02, 0.02, 3)
        x0[3:6] += rng.uniform(-0.05, 0.05, 3)
        x0[6:6 + nj] = np.clip(x0[6:6 + nj] + rng.uniform(-0.1, 0.1, nj), lo + 0.05, hi - 0.05)
        x0[6 + nj:] += rng.uniform(-0.2, 0.2, 6 + nj)
        cmd = [rng.uniform(-0.5, 1.0), rng.uniform(-0.3, 0.3), model["reference"]["defaultBaseHeight"], rng.uniform(-0.5, 0.5)]
        g = gaits[i % len(gaits)]
        start = None
        if random_phase:
            period = model["gaits"][g]["switchingTimes"][-1] if g != "stance" else 1.0
            start = -rng.uniform(0.0, period)
        insts.append(references.build_instance(model, x0, t0=0.0, horizon=horizon, gait=g, gait_start=start, cmd=cmd))
        insts[-1]["cmd"], insts[-1]["gait"] = cmd, g
    return insts


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def oracle_batch_solve(model, insts, settings, threads):
    """The CHECKER oracle (dense dual-number Jacobians, oracle/wb_problem.hpp) on the host cores, one instance per worker thread.  Slow by
    construction; only used by --check-oracle."""
    import ctypes as C

    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_lib as orc

    from wb_humanoid_mpc_b200.solver import stack_instances

    L = orc.lib()
    b = stack_instances(insts)
    desc = abi.model_desc(model)
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    u8 = lambda a: np.ascontiguousarray(a, dtype=np.uint8)
    x, u = f(b["x_init"]).copy(), f(b["u_init"]).copy()
    arrs = [f(b["t_nodes"]), u8(b["node_event"]), f(b["x0"]), x, u, u8(b["contact_flags"]), f(b["swing_ref"]), f(b["impact_factor"]), f(b["arm_phase"]),
            f(b["x_ref"])]
    u8p = C.POINTER(C.c_uint8)
    t0 = time.perf_counter()
    rc = L.orc_wb_sqp_batch(C.byref(desc), C.c_int(len(insts)), C.c_int(threads), C.c_int(b["t_nodes"].shape[1]), orc._p(arrs[0]),
                            arrs[1].ctypes.data_as(u8p), orc._p(arrs[2]), orc._p(arrs[3]), orc._p(arrs[4]), arrs[5].ctypes.data_as(u8p),
                            orc._p(arrs[6]), orc._p(arrs[7]), orc._p(arrs[8]), orc._p(arrs[9]), C.byref(settings))
    dt = time.perf_counter() - t0
    assert rc == 0, rc
    return dt, x, u


def effective_cores():
    """host threads this process can actually keep busy: the CPU affinity mask, capped by the container's cgroup CPU quota (a box may expose
    128 hardware threads and grant 12 CPUs worth of time; oversubscribing the quota only adds context switches)"""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    quota = None
    try:
        q, per = Path("/sys/fs/cgroup/cpu.max").read_text().split()[:2]
        if q != "max":
            quota = float(q) / float(per)
    except Exception:
        try:
            q = float(Path("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read_text())
            per = float(Path("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read_text())
            if q > 0:
                quota = q / per
        except Exception:
            pass
    eff = n if quota is None else max(1, min(n, int(np.ceil(quota))))
    return eff, n, quota


def cpu_port_solve(model, insts, settings, threads, node_threads=1):
    """The TIMED CPU arm: the fast restatement of the reference's CPU path (oracle/fast/wb_fast.cu: analytic Jacobians per node, dense
    sequential Riccati, value-only roll-outs), pinned on the checker oracle by tests/test_oracle_fast.py.  -> oracle_lib.fast_wb_sqp_batch dict"""
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_lib as orc

    from wb_humanoid_mpc_b200.solver import stack_instances

    return orc.fast_wb_sqp_batch(model, stack_instances(insts), settings, threads=threads, node_threads=node_threads)


def calibrate_threads(model, sample, settings, cores):
    """The box may grant less CPU time than its hardware-thread count suggests (container quotas that /sys does not show): the worker-thread
    count of the CPU arm is calibrated -- 2 short repetitions each at n, n/2, n/4, n/8 threads -- and the best one is used and REPORTED."""
    best, best_thr = cores, 0.0
    t_try = cores
    while t_try >= 1:
        tt = min(cpu_port_solve(model, sample, settings, t_try)["seconds"] for _ in range(2))
        if len(sample) / tt > best_thr * 1.03:
            best, best_thr = t_try, len(sample) / tt
        if t_try == 1 or t_try <= cores // 8:
            break
        t_try = max(1, t_try // 2)
    return best


def cpu_baseline_rows(model, insts, settings, cores, reps, warm=3):
    """CPU-B "host throughput" (one instance per thread on all host cores) and CPU-A "reference-like latency"
    (one instance, its shooting nodes on 4 threads like task.info nThreads 4, sequential Riccati); `warm` warm-up + `reps` timed repetitions,
    median and p95.  Returns (row dict for the JSON line, the solutions of the CPU-B sample)."""
    sample = insts[: max(1, min(len(insts), 2 * cores))]
    for _ in range(warm):
        cpu_port_solve(model, sample[: max(1, cores // 2)], settings, cores)
    best = calibrate_threads(model, sample, settings, cores)
    hw_threads, cores = cores, best
    tb, out = [], None
    for _ in range(reps):
        out = cpu_port_solve(model, sample, settings, cores)
        tb.append(out["seconds"])
    tb = np.array(tb)
    thr = len(sample) / tb
    ta = []
    for i in range(warm + reps):
        o = cpu_port_solve(model, insts[:1], settings, 1, node_threads=4)
        if i >= warm:
            ta.append(o["seconds"])
    ta = np.array(ta)
    st = out["stage_s"] / len(sample)
    row = {"value": float(np.median(thr)), "unit": "solves/s", "cores": cores, "kind": "port",
           "host": {"hardware_threads": hw_threads, "affinity_cpus": effective_cores()[1], "cgroup_cpu_quota": effective_cores()[2],
                    "worker_threads": "calibrated: best of n, n/2, n/4, n/8"},
           "what": "fast CPU restatement of the reference path (oracle/fast/wb_fast.cu: analytic per-node Jacobians on host threads, dense sequential Riccati, "
                   "value-only roll-outs), NOT the ocs2+HPIPM binary (not buildable here, DESIGN.md section 2); pinned on the checker oracle by tests/test_oracle_fast.py",
           "sample": f"CPU-B: {len(sample)} instances of the workload per repetition, one instance per thread on {cores} threads, {warm} warm-up + {reps} "
                     f"repetitions ({float(tb.sum()):.1f} s)",
           "p95_low": float(np.quantile(thr, 0.05)), "reps": int(reps),
           "core_ms_per_solve": {"lq": 1e3 * float(st[0]), "qp": 1e3 * float(st[1]), "linesearch": 1e3 * float(st[2])},
           "cpu_a_latency": {"what": "CPU-A: 1 instance, shooting nodes on 4 threads (task.info nThreads 4), sequential Riccati", "median_ms": 1e3 * float(np.median(ta)),
                             "p95_ms": 1e3 * float(np.quantile(ta, 0.95)), "solves_per_s": float(1.0 / np.median(ta)), "threads": 4, "reps": int(reps)}}
    return row, out, sample


def run_mixed(args, model, settings, rank, world, local_rank, workload, cores):
    """BASELINE configs[4]: the mixed contact-schedule sweep.  Instances with different event counts have different numbers of shooting nodes;
    they are grouped by node count, one library handle per group, the groups solved concurrently (one host thread and CUDA stream per group).
    `value`: device-resident (reset + solve of every group per step); `e2e`: upload from pinned host memory + solve + download of every group per
    step.  Timed by wall clock between device synchronisations (CUDA events on one stream cannot bracket multi-stream work), max over ranks."""
    from concurrent.futures import ThreadPoolExecutor

    import torch

    from wb_humanoid_mpc_b200.solver import B200SqpSolver, stack_instances

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    insts = build_batch(model, args.batch, rank, args.horizon, ["stance", "walk", "slow_walk"], random_phase=True)
    by_n = {}
    for i in insts:
        by_n.setdefault(len(i["t_nodes"]), []).append(i)
    groups = []
    for n, gi in sorted(by_n.items()):
        b = stack_instances(gi)
        pinned = {k: torch.from_numpy(np.ascontiguousarray(v if v.dtype == np.uint8 else v.astype(np.float64))).pin_memory().numpy() for k, v in b.items()}
        sv = B200SqpSolver(model, settings, device=local_rank)
        sv.upload(pinned)
        groups.append({"n": n, "B": len(gi), "solver": sv, "pinned": pinned, "stream": torch.cuda.Stream()})
    pool = ThreadPoolExecutor(max_workers=len(groups))

    def each(fn):
        def work(g):
            torch.cuda.set_device(local_rank)
            fn(g)
        list(pool.map(work, groups))

    def resident(g):
        g["solver"].reset()
        g["solver"].solve(g["stream"].cuda_stream)

    def e2e_step(g):
        g["solver"].upload(g["pinned"])
        g["solver"].solve(g["stream"].cuda_stream)
        g["out"] = g["solver"].primal_solution()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            each(fn)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        barrier()
        return dt

    for _ in range(args.warmup):
        each(resident)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dev_s = timed(resident, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = sum(g["solver"].launch_count() for g in groups) * args.steps
    stage = {str(g["n"]): dict(zip(["lq", "qp", "linesearch", "lq_projection_share"], [float(v) for v in g["solver"].benchmarks()])) for g in groups}
    each(e2e_step)
    e2e_s = timed(e2e_step, args.steps)
    assert all(not g["out"]["status"].any() for g in groups)
    total = args.batch * world * args.steps
    h2d = sum(sum(v.nbytes for v in g["pinned"].values()) for g in groups)
    d2h = sum(g["B"] * g["n"] * 58 * 8 + g["B"] * (g["n"] - 1) * 35 * 8 + g["B"] * settings.sqp_iteration * 128 + g["B"] * 8 for g in groups)
    line = {"metric": METRIC, "value": total / dev_s, "unit": "solves/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "node_count_groups": {str(g["n"]): g["B"] for g in groups}, "batch_per_gpu": args.batch,
                       "timing": "wall clock between device synchronisations (several CUDA streams), max over ranks",
                       "l2": "stage records (GBs per GPU) exceed the 126 MB L2; no flush needed"},
            "e2e": {"value": total / e2e_s, "unit": "solves/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "mode": "every group: upload from pinned host memory -> solve -> download, groups concurrent on their own streams"},
            "gpu_launches": int(launches), "stage_ms_by_node_count": stage, "clocks": clocks,
            "roofline": None}
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c3", choices=["c3", "c4", "c5"],
                    help="BASELINE.json configs[2..4]: c3 = walk, 256 instances per GPU (the metric's configuration, default); c4 = walk, 1024 per GPU "
                         "(8192 over 8 GPUs); c5 = mixed contact-schedule sweep stance/walk/slow_walk with random gait phase, 1024 per GPU (4096 over 4)")
    ap.add_argument("--batch", type=int, default=0, help="MPC instances per GPU (default: what --config says)")
    ap.add_argument("--horizon", type=float, default=3.5, help="seconds; 3.5 s at dt = 0.035 s gives N = 100 intervals (+ event nodes)")
    ap.add_argument("--gait", default="walk")
    ap.add_argument("--cpu-sample", type=int, default=0, help="instances in the CPU-baseline sample (0 = 2 x cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-reps", type=int, default=20, help="timed repetitions of the CPU baseline rows (after 3 warm-up)")
    ap.add_argument("--check-oracle", action="store_true", help="additionally cross-check the GPU solution against the (slow) checker oracle")
    ap.add_argument("--sqp-iteration", type=int, default=1, help="sqpIteration (1 = the shipped real-time iteration; 10 = the secondary number)")
    ap.add_argument("--global-step", action="store_true", help="one line-search step per iteration for the whole multi-GPU batch (NCCL)")
    args = ap.parse_args()
    if args.batch <= 0:
        args.batch = 256 if args.config == "c3" else 1024

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    model = model_loader.load_packaged_model()
    settings = abi.default_settings(model, sqp_iteration=args.sqp_iteration, global_step=int(args.global_step))
    cores, affinity_cpus, cpu_quota = effective_cores()
    n_int = int(round(args.horizon / model["sqp"]["dt"]))
    workload = f"G1 whole-body MPC (nx=58, nu=35), dt=0.035 s x {n_int} intervals, gait={args.gait}, batch={args.batch}/GPU, sqpIteration={args.sqp_iteration}, cold start" + (", global line-search step" if args.global_step else "")

    if args.config == "c5":
        workload = (f"G1 whole-body MPC (nx=58, nu=35), dt=0.035 s x {n_int} intervals, mixed contact-schedule sweep (stance / walk / slow_walk, random gait "
                    f"phase), batch={args.batch}/GPU, sqpIteration={args.sqp_iteration}, cold start")
        if args.impl != "reference":
            return run_mixed(args, model, settings, rank, world, local_rank, workload, cores)

    if args.impl == "reference":
        # the reference's own CPU implementation of the path cannot be built here (no Eigen/Pinocchio/HPIPM): the arm
        # times the fast CPU restatement of the same algorithm (kind "port") on all host cores, rank 0 only.
        if rank != 0:
            return
        sample = args.cpu_sample or 2 * cores
        if args.config == "c5":   # the mixed sweep: instances grouped by node count, the groups one after the other (each on all cores)
            insts = build_batch(model, sample, 0, args.horizon, ["stance", "walk", "slow_walk"], random_phase=True)
            by_n = {}
            for i in insts:
                by_n.setdefault(len(i["t_nodes"]), []).append(i)
            groups = list(by_n.values())
        else:
            insts = build_batch(model, sample, 0, args.horizon, [args.gait])
            groups = [insts]
        for _ in range(args.warmup):
            cpu_port_solve(model, groups[0][: max(1, cores // 2)], settings, cores)
        hw_threads = cores
        cores = calibrate_threads(model, max(groups, key=len), settings, cores)
        times, stage = [], np.zeros(3)
        for _ in range(args.steps):
            tstep = 0.0
            for g in groups:
                o = cpu_port_solve(model, g, settings, cores)
                tstep += o["seconds"]
                stage += o["stage_s"]
            times.append(tstep)
        total = sum(times)
        val = sample * args.steps / total
        ta = [cpu_port_solve(model, insts[:1], settings, 1, node_threads=4)["seconds"] for _ in range(8)][3:]
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": {"workload": workload, "n_nodes": len(insts[0]["t_nodes"]), "sample_instances_per_step": sample},
                "cpu_baseline": {"value": val, "unit": "solves/s", "cores": cores, "kind": "port",
                                 "host": {"hardware_threads": hw_threads, "affinity_cpus": affinity_cpus, "cgroup_cpu_quota": cpu_quota,
                                          "worker_threads": "calibrated: best of n, n/2, n/4, n/8"},
                                 "what": "fast CPU restatement of the reference path (oracle/fast/wb_fast.cu), NOT the ocs2+HPIPM binary (not buildable here)",
                                 "sample": f"{sample} instances per step, one instance per thread, {cores} threads",
                                 "median_step_solves_per_s": float(np.median(sample / np.array(times))),
                                 "core_ms_per_solve": dict(zip(["lq", "qp", "linesearch"], (1e3 * stage / (sample * args.steps)).tolist())),
                                 "cpu_a_latency_ms": 1e3 * float(np.median(ta))},
                "e2e": {"value": val, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch

    from wb_humanoid_mpc_b200.solver import B200SqpSolver, stack_instances

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 arm has no CPU fallback (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    insts = build_batch(model, args.batch, rank, args.horizon, [args.gait])
    batch = stack_instances(insts)
    n_nodes = batch["t_nodes"].shape[1]
    solver = B200SqpSolver(model, settings, device=local_rank)
    if args.global_step:
        solver.enable_global_step()

    # pinned host staging buffers for the end-to-end path
    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t.numpy()

    pinned = {k: pin(v if v.dtype == np.uint8 else v.astype(np.float64)) for k, v in batch.items()}
    h2d = sum(v.nbytes for v in pinned.values())
    B, nx, nu = args.batch, model["nx"], model["nu"]
    d2h = B * n_nodes * nx * 8 + B * (n_nodes - 1) * nu * 8 + B * settings.sqp_iteration * 128 + B * 8

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(seconds):
        if dist is None:
            return seconds
        t = torch.tensor([seconds], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput (value) --------------------------------------------------------------------------------
    solver.upload(pinned)
    for _ in range(args.warmup):
        solver.reset()
        solver.solve()
    stage_acc = np.zeros(4)
    launches = 0
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        solver.reset()
        solver.solve()   # synchronises internally at the end of the line search (host reads the pending-instance counter)
        stage_acc += np.array(solver.benchmarks())
        launches += solver.launch_count()
    ev1.record()
    torch.cuda.synchronize()
    dev_s = max_over_ranks(ev0.elapsed_time(ev1) * 1e-3)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    sol = solver.primal_solution()
    assert not sol["status"].any()
    alphas = sol["log"][:, 0, 8]

    # ---- end-to-end through the C ABI with host buffers (e2e) --------------------------------------------------------------------
    for _ in range(min(args.warmup, 2)):
        solver.upload(pinned)
        solver.solve()
        solver.primal_solution()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        solver.upload(pinned)     # H2D of every per-instance input from pinned host memory
        solver.solve()
        solver.primal_solution()  # D2H of the primal solution + iteration log
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()

    total_solves = args.batch * world * args.steps
    value = total_solves / dev_s
    e2e_serial = total_solves / e2e_s

    # ---- the same, double-buffered: two handles on two CUDA streams driven by two host threads, each step = {upload from pinned host
    # memory, solve, download into pinned host memory} of one full batch; the copies (and the latency-bound Riccati sweep) of one batch
    # overlap the LQ approximation of the other.  Every step still moves its own inputs and results; this is the serving pattern.
    e2e = e2e_serial
    e2e_pipe = None
    solvers = None
    if not args.global_step and args.steps >= 2:
        import threading

        try:   # the second handle doubles the device footprint (4.1 GB per 256 instances): very large per-GPU batches fall back to serial
            solvers = [solver, B200SqpSolver(model, settings, device=local_rank)]
            solvers[1].upload(pinned)
        except Exception as e:
            print(f"bench.py: double-buffered e2e unavailable ({e!r}); reporting the serial number", file=sys.stderr)
            solvers = None
        if dist is not None:   # the double-buffered leg contains barriers: every rank runs it or none does
            okt = torch.tensor([1 if solvers is not None else 0], dtype=torch.int32, device="cuda")
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
            if int(okt.item()) == 0:
                solvers = None
    if solvers is not None:
        streams = [torch.cuda.Stream(), torch.cuda.Stream()]
        outs = [{"x": pin(np.zeros((B, n_nodes, nx))), "u": pin(np.zeros((B, n_nodes - 1, nu)))} for _ in range(2)]
        results = [None, None]

        def worker(i, n):
            torch.cuda.set_device(local_rank)
            for _ in range(n):
                solvers[i].upload(pinned)
                solvers[i].solve(streams[i].cuda_stream)
                results[i] = solvers[i].primal_solution(out=outs[i])

        def run_pipe(n_each):
            th = [threading.Thread(target=worker, args=(i, n_each[i])) for i in range(2)]
            for t in th:
                t.start()
            for t in th:
                t.join()
            torch.cuda.synchronize()

        run_pipe([1, 1])
        barrier()
        t0 = time.perf_counter()
        run_pipe([(args.steps + 1) // 2, args.steps // 2])
        pipe_s = max_over_ranks(time.perf_counter() - t0)
        barrier()
        assert all(r is not None and not r["status"].any() for r in results)
        pipe_dx = max(float(np.abs(r["x"] - sol["x"]).max()) for r in results)
        e2e_pipe = {"value": total_solves / pipe_s, "handles_in_flight": 2, "max_abs_diff_x_vs_serial": pipe_dx}
        e2e = e2e_pipe["value"]
        del solvers[1]

    # ---- the same through the C++ host layer (b200sqp::host::SqpSolver::run, the mirror of ocs2::SqpSolver::run): per-instance reference
    # managers, time grids and cold-start initial guesses are built on host threads inside the timed region, then upload + solve + download
    host_api = None
    try:
        from wb_humanoid_mpc_b200 import host_lib

        hm = host_lib.HostModel()
        hs = host_lib.HostSqpSolver(hm, settings, args.batch, device=local_rank)
        x0s = np.array([i["x0"] for i in insts])
        for b, i in enumerate(insts):
            hs.set_gait(b, i["gait"], 0.0, 3 * args.horizon)
            hs.set_command(b, 0.0, i["x0"], i["cmd"], args.horizon)
        for _ in range(min(args.warmup, 2)):
            hs.reset()
            hs.run(0.0, x0s, args.horizon)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            hs.reset()   # cold start every step, like the other legs
            hs.run(0.0, x0s, args.horizon)
        host_s = max_over_ranks(time.perf_counter() - t0)
        barrier()
        dx = max(float(np.abs(hs.primal_solution(b)["x"] - sol["x"][b]).max()) for b in range(0, args.batch, max(1, args.batch // 8)))
        hb = hs.benchmarks()
        host_api = {"value": total_solves / host_s, "unit": "solves/s", "call": "b200sqp::host::SqpSolver::run (C++ host layer, instances built on host threads)",
                    "max_abs_diff_x_vs_c_abi_path": dx,
                    "last_run_ms": dict(zip(["pre_run", "pack", "upload", "solve", "download", "unpack"], [round(float(v), 3) for v in hb[4:10]]))}
        # double-buffered like e2e above: two SqpSolver objects on two host threads (each owns its handles and CUDA stream)
        if args.steps >= 2:
            import threading

            hs2 = host_lib.HostSqpSolver(hm, settings, args.batch, device=local_rank)
            hs.set_exclusive_solve(True)    # one solve on the GPU at a time; the other object builds / unpacks its batch meanwhile
            hs2.set_exclusive_solve(True)
            for b, i in enumerate(insts):
                hs2.set_gait(b, i["gait"], 0.0, 3 * args.horizon)
                hs2.set_command(b, 0.0, i["x0"], i["cmd"], args.horizon)

            def hworker(sv, n):
                for _ in range(n):
                    sv.reset()
                    sv.run(0.0, x0s, args.horizon)

            def hrun(n_each):
                th = [threading.Thread(target=hworker, args=(sv, n)) for sv, n in zip((hs, hs2), n_each)]
                for t in th:
                    t.start()
                for t in th:
                    t.join()

            hrun([1, 1])
            barrier()
            t0 = time.perf_counter()
            hrun([(args.steps + 1) // 2, args.steps // 2])
            hp_s = max_over_ranks(time.perf_counter() - t0)
            barrier()
            host_api["serial"] = {"value": host_api["value"], "unit": "solves/s"}
            host_api["value"] = total_solves / hp_s
            host_api["mode"] = "double-buffered: 2 SqpSolver objects on 2 host threads, setExclusiveSolve (device phases take turns)"
            host_api["max_abs_diff_x_second_solver"] = max(float(np.abs(hs2.primal_solution(b)["x"] - sol["x"][b]).max())
                                                           for b in range(0, args.batch, max(1, args.batch // 8)))
            hs2.close()
        hs.close()
        hm.close()
    except Exception as e:   # the host layer is optional for the bench line
        host_api = {"unavailable": repr(e)}

    # ---- the same with the DEVICE-SIDE INSTANCE BUILDER (b200sqp_build_instances): per step the host hands over x0, gait id, gait start and the
    # velocity command of every instance (~0.5 kB each, pinned), the GPU builds mode schedule / swing references / targets / time grid / initial
    # guess, solves, and the primal solution comes back -- the end-to-end path without the 150 kB per instance of per-node host arrays
    device_builder = None
    try:
        bx0 = pin(np.array([i["x0"] for i in insts]))
        bcmd = pin(np.array([i["cmd"] for i in insts], dtype=np.float64))
        bstart = pin(np.zeros(B))
        bgait = [i["gait"] for i in insts]
        sb = B200SqpSolver(model, settings, device=local_rank)
        for _ in range(min(args.warmup, 2)):
            sb.build_instances(0.0, args.horizon, bx0, bgait, bstart, bcmd)
            sb.solve()
            rb = sb.primal_solution()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            sb.build_instances(0.0, args.horizon, bx0, bgait, bstart, bcmd)
            sb.solve()
            rb = sb.primal_solution()
        torch.cuda.synchronize()
        b_s = max_over_ranks(time.perf_counter() - t0)
        barrier()
        device_builder = {"value": total_solves / b_s, "unit": "solves/s", "h2d_bytes_per_step": int(bx0.nbytes + bcmd.nbytes + bstart.nbytes + 4 * B),
                          "d2h_bytes_per_step": int(d2h), "call": "b200sqp_build_instances + b200sqp_solve + b200sqp_download (serial, one handle)",
                          "max_abs_diff_x_vs_upload_path": float(np.abs(rb["x"] - sol["x"]).max())}
        if args.steps >= 2 and not args.global_step:
            # double-buffered like `e2e`: two handles on two streams, two host threads; the download of one batch overlaps the solve of the other
            import threading

            sbs = [sb, B200SqpSolver(model, settings, device=local_rank)]
            bstreams = [torch.cuda.Stream(), torch.cuda.Stream()]
            bouts = [{"x": pin(np.zeros((B, n_nodes, nx))), "u": pin(np.zeros((B, n_nodes - 1, nu)))} for _ in range(2)]

            def bworker(i, n):
                torch.cuda.set_device(local_rank)
                for _ in range(n):
                    sbs[i].build_instances(0.0, args.horizon, bx0, bgait, bstart, bcmd)
                    sbs[i].solve(bstreams[i].cuda_stream, wait=False)
                    sbs[i].primal_solution(out=bouts[i])

            def brun(n_each):
                th = [threading.Thread(target=bworker, args=(i, n_each[i])) for i in range(2)]
                for t in th:
                    t.start()
                for t in th:
                    t.join()
                torch.cuda.synchronize()

            brun([1, 1])
            barrier()
            t0 = time.perf_counter()
            brun([(args.steps + 1) // 2, args.steps // 2])
            bp_s = max_over_ranks(time.perf_counter() - t0)
            barrier()
            device_builder["serial"] = {"value": device_builder["value"], "unit": "solves/s"}
            device_builder["value"] = total_solves / bp_s
            device_builder["call"] = "b200sqp_build_instances + b200sqp_solve + b200sqp_download, double-buffered (2 handles, 2 streams, 2 host threads)"
            sbs[1].close()
        sb.close()
    except Exception as e:
        device_builder = {"unavailable": repr(e)}

    # ---- per-kernel roofline -------------------------------------------------------------------------------------------
    # Device ms per stage are CUDA events recorded by the library on the launching stream (b200sqp_get_stage_times): ms[0] = K1a + K1b,
    # ms[3] = lu_kernel + the two K1b kernels, ms[1] = K2 (+ remap), ms[2] = the line search = n_ls x (K3 + accept).
    stage_ms = stage_acc / args.steps
    N = n_nodes - 1
    nut = 23
    # K3 launches that do work, in units of a full-batch launch: the library enqueues the whole back-tracking ladder (asynchronous solve) and the
    # thread blocks of finished instances return at once; an instance accepted at alpha = decay^t ran t + 1 trials
    trials = np.where(alphas > 0, np.round(np.log(np.maximum(alphas, 1e-300)) / np.log(settings.alpha_decay)) + 1, 14)
    n_ls = max(1.0, float(np.mean(trials)))
    k_ms = {"lq_dyn_kernel (K1a)": stage_ms[0] - stage_ms[3], "lu + lq_projdyn + lq_proj kernels (K1b)": stage_ms[3], "riccati_bwd + riccati_fwd kernels (K2)": stage_ms[1],
            "rollout_kernel (K3)": stage_ms[2] / n_ls}
    k_share = {"lq_dyn_kernel (K1a)": k_ms["lq_dyn_kernel (K1a)"], "lu + lq_projdyn + lq_proj kernels (K1b)": stage_ms[3], "riccati_bwd + riccati_fwd kernels (K2)": stage_ms[1],
               "rollout_kernel (K3)": stage_ms[2]}
    # algorithmic bytes per launch: what each kernel must read + write given the kernel split (doubles x 8)
    rec = 8 * (58 * 58 + 58 * nut + 58 + 58 * 58 + nut * 58 + nut * nut + 58 + nut)            # projected stage record A B b Q S R q r
    proj = 8 * (35 * nut + 35 * 58 + 35)                                                    # Pu Px u0 (read by the remap)
    node_in = 8 * (58 + 35 + 58 + 58 + 6 + 2 + 1 + 1) + 3                                   # x u x+ xref swing impact arm t, flags
    swing_rows = 15.0 * float((1 - batch["contact_flags"][:, :-1, :].astype(np.float64)).sum()) / (B * N)   # mean dense cost rows / node
    mid = 8 * (12 * 93 + 58 + 14 * 93 + 14 + 93 + 93 + 18 + 24 * 27 + 6 + swing_rows * 93)   # K1a -> K1b record (struct Mid)
    lu = 8 * (14 * 35) + 2 * (8 * 15 * 35 + 4 * 52) + 2 * 8 * (14 * 59 + 14 * 23) + 8 * (12 * 93 + 58 + 14 * 94)   # lu_kernel: D in, factors + permutations out (and back in); part 1 -> part 2: [X | x0], K out and back in; part 1 re-reads A/B rows, b, C, D, e
    alg = {"lq_dyn_kernel (K1a)": B * N * (node_in + mid), "lu + lq_projdyn + lq_proj kernels (K1b)": B * N * (mid + rec + proj + lu),
           # backward sweep: the record in, K~ k out; forward sweep (own kernel): A, b, B~, K~, k in, dx, du~ out
           "riccati_bwd + riccati_fwd kernels (K2)": B * N * (rec + 8 * (nut * 58 + nut) + 8 * (58 * 58 + 58 + 58 * nut + nut * 58 + nut) + 8 * (58 + nut)),
           "rollout_kernel (K3)": B * N * (8 * (3 * 58 + 35 + 58 + 35) + 32)}
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    fp64_peak = 40.0   # TFLOP/s, nominal B200 fp64 (vector and DMMA); MEASURED_PEAKS.json carries bf16 only
    short = {"lq_dyn_kernel (K1a)": ["lqa"], "lu + lq_projdyn + lq_proj kernels (K1b)": ["lqb", "lqp", "lu"], "riccati_bwd + riccati_fwd kernels (K2)": ["ricb", "ricf"],
             "rollout_kernel (K3)": ["ro"]}
    kernels = {}
    for name, ms in k_ms.items():
        e = {"ms_per_launch": ms, "ms_per_step": k_share[name], "algorithmic_bytes_per_launch": alg[name],
             "hbm_gbs": alg[name] / (ms * 1e-3) / 1e9, "hbm_frac": alg[name] / (ms * 1e-3) / 1e9 / peak}
        try:   # counters of the latest committed `ncu --set full` capture(s) of this kernel (the captured batch is recorded in the file)
            e["traffic"], e["ncu_capture"], tflop = 0.0, [], 0.0
            for sh in short[name]:
                raws = sorted((ROOT / "profiles").glob(f"ncu_{sh}_*_raw.json"))
                if not raws:
                    continue   # no capture of this kernel committed (yet)
                raw = json.loads(raws[-1].read_text())
                scale = B / float(raw.get("batch", 64))   # the r2 captures are taken at the benchmarked batch (256): scale 1
                e["ncu_capture"].append(raws[-1].name)
                e["traffic"] += (float(raw["dram_bytes_read_B"]) + float(raw["dram_bytes_write_B"])) * scale
                tflop += float(raw.get("sm__ops_path_tensor_src_fp64.sum", 0.0)) * scale
                if sh == short[name][0]:
                    e["ncu_pipe_pct"] = {"fp64": float(raw["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]),
                                         "dmma": float(raw["sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active"])}
            e["dmma_tflops"] = tflop / (ms * 1e-3) / 1e12
            e["dmma_frac_of_nominal_fp64"] = e["dmma_tflops"] / fp64_peak
        except Exception:
            e["traffic"] = None
        kernels[name] = e
    dom = max(k_share, key=k_share.get)
    kd = kernels[dom]
    roofline = {"kernel": dom, "bound": "hbm", "achieved": kd["hbm_gbs"], "peak": peak, "unit": "GB/s", "frac": kd["hbm_frac"],
                "traffic": kd.get("traffic"), "peak_source": "MEASURED_PEAKS.json (measured)" if peaks else "fallback 6.65 TB/s",
                "algorithmic_bytes_per_launch": kd["algorithmic_bytes_per_launch"],
                "stage_ms": {"lq": stage_ms[0], "lq_projection_share": stage_ms[3], "qp": stage_ms[1], "linesearch": stage_ms[2]},
                "kernels": kernels, "fp64_peak_tflops_nominal": fp64_peak,
                "note": "north_star asks for the HBM fraction; none of the kernels is HBM bound (DESIGN.md §6): K2's contractions run at the "
                        "DMMA rate shown, K1a/K1b/K3 are barrier/latency bound.  Times are CUDA events on the launching stream."}
    rec_gb = B * N * (rec + proj + mid) / 1e9

    line = {"metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload, "n_nodes": int(n_nodes), "batch_per_gpu": args.batch, "l2": "stage records (%.1f GB/GPU) exceed the 126 MB L2; no flush needed" % rec_gb,
                       "accepted_step_sizes": {str(a): int((alphas == a).sum()) for a in np.unique(alphas)}},
            "e2e": {"value": e2e, "unit": "solves/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "mode": "double-buffered: 2 handles in flight on 2 CUDA streams, every step uploads its inputs from and downloads its results to "
                            "pinned host memory" if e2e_pipe else "serial upload -> solve -> download",
                    "pipelined": e2e_pipe, "serial": {"value": e2e_serial, "unit": "solves/s"}, "host_api": host_api, "device_builder": device_builder},
            "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        row, cpu_out, cpu_sample = cpu_baseline_rows(model, insts, settings, cores, reps=args.cpu_reps)
        line["cpu_baseline"] = row
        # parity spot check on the benchmark inputs: the CPU arm and the GPU arm solved the same instances.  Instances whose complete-pivoting
        # LU sees the same pivot order agree to ~1e-11; a pivot tie resolved differently changes the basis of the projected QP and the two
        # (equally valid) answers then differ by the conditioning of the problem, ~1e-6 relative -- asserted at 1e-5.
        k = len(cpu_sample)
        dxs = np.abs(sol["x"][:k] - cpu_out["x"][:k]).reshape(k, -1).max(1) / np.abs(cpu_out["x"][:k]).reshape(k, -1).max(1)
        row["rel_diff_x_vs_gpu"] = {"max": float(dxs.max()), "median": float(np.median(dxs)), "instances": int(k)}
        assert dxs.max() < 1e-5, f"GPU and CPU arms disagree on the benchmark inputs: {dxs.max():.3e}"
        assert np.array_equal(sol["log"][:k, 0, 8], cpu_out["log"][:k, 0, 8]), "GPU and CPU arms accepted different step sizes"
        if args.check_oracle:
            ko = min(k, cores)
            dt, xo, uo = oracle_batch_solve(model, cpu_sample[:ko], settings, cores)
            eo = np.abs(sol["x"][:ko] - xo).reshape(ko, -1).max(1) / np.abs(xo).reshape(ko, -1).max(1)
            row["checker_oracle"] = {"seconds": dt, "instances": ko, "rel_diff_x_vs_gpu_max": float(eo.max()), "rel_diff_x_vs_gpu_median": float(np.median(eo))}
            assert eo.max() < 1e-5
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
