"""Per-phase cycle table of the phase-scheduled kernels (development aid).

  python tools/phase_clock.py --build            # here: cross-compile the instrumented library (tools/_clk/, git-ignored)
  gpurun -- python tools/phase_clock.py --node 20 --batch 256   # on the GPU box: one SQP iteration, print the tables

The instrumented build (-DB200SQP_PHASE_CLOCK) records clock64() of thread 0 of ONE CTA after every phase barrier; the CTA runs inside the
full grid, so the cycles include the contention of its co-resident CTAs.  Not a product path and never a bench number.
"""
import argparse
import ctypes as C
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
os.environ["B200SQP_LIB"] = str(ROOT / "tools" / "_clk" / "libb200sqp_clk.so")
os.environ["B200SQP_NVCC_EXTRA"] = "-DB200SQP_PHASE_CLOCK"

ap = argparse.ArgumentParser()
ap.add_argument("--build", action="store_true")
ap.add_argument("--node", type=int, default=20)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--top", type=int, default=200)
args = ap.parse_args()

from wb_humanoid_mpc_b200 import lib  # noqa: E402

if args.build:
    lib.build(force=True)
    print("built", lib.SO)
    sys.exit(0)

import numpy as np  # noqa: E402

import bench  # noqa: E402
from wb_humanoid_mpc_b200 import abi, model_loader  # noqa: E402
from wb_humanoid_mpc_b200.solver import B200SqpSolver, stack_instances  # noqa: E402

model = model_loader.load_packaged_model()
settings = abi.default_settings(model, sqp_iteration=1)
insts = bench.build_batch(model, args.batch, 0, 3.5, ["walk"])
batch = stack_instances(insts)
solver = B200SqpSolver(model, settings, device=0)
L = lib.lib()
L.b200sqp_debug_phase_clocks.restype = C.c_int
L.b200sqp_debug_phase_clocks.argtypes = [C.c_int, C.POINTER(C.c_longlong), C.c_int, C.c_int]
solver.upload(batch)
node = args.node
print("node", node, "contact", batch["contact_flags"][0, node], "event", batch["node_event"][0, node])
L.b200sqp_debug_phase_clocks(0, None, 0, node)
for _ in range(2):
    solver.reset()
    solver.solve()
files = {0: "wb_node_a.inc", 1: "wb_node_b2.inc", 2: "wb_rollout_body.inc", 4: "wb_node_b1.inc"}
cost_src = (ROOT / "wb_humanoid_mpc_b200" / "csrc" / "wb_node_b_cost.inc").read_text().split("\n")
for kid, name in [(0, "K1a lq_dyn_kernel"), (4, "K1b part 1 lq_projdyn_kernel"), (1, "K1b part 2 lq_proj_kernel"), (2, "K3 rollout_kernel")]:
    buf = (C.c_longlong * (2 * 512))()
    n = L.b200sqp_debug_phase_clocks(kid, buf, 512, -1)
    a = np.array(buf[:2 * n], dtype=np.int64).reshape(n, 2)
    src = (ROOT / "wb_humanoid_mpc_b200" / "csrc" / files[kid]).read_text().split("\n")
    tot = a[-1, 1] - a[0, 1]
    print(f"\n== {name}: {n - 1} phases, {tot} cycles")
    rows = []
    for i in range(1, n):
        rows.append((int(a[i, 1] - a[i - 1, 1]), int(a[i, 0]), i))
    # aggregate by source line (RK stages repeat the same lines)
    agg = {}
    for cyc, line, i in rows:
        c, k = agg.get(line, (0, 0))
        agg[line] = (c + cyc, k + 1)
    for line, (cyc, k) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:args.top]:
        text = (cost_src[line - 1000].strip()[:110] if line > 1000 else src[line - 1].strip()[:110]) if 0 < line <= 1000 + len(cost_src) else "?"
        print(f"  {cyc:8d} {100.0 * cyc / tot:5.1f}%  x{k:<2d} L{line:<4d} {text}")

buf = (C.c_longlong * 64)()
n = L.b200sqp_debug_phase_clocks(3, buf, 32, -1)
a = np.array(buf[:64], dtype=np.int64).reshape(32, 2)
legacy = os.environ.get("B200SQP_K2_LEGACY") == "1"
src = (ROOT / "wb_humanoid_mpc_b200" / "csrc" / ("riccati.cuh" if legacy else "riccati_wb.cuh")).read_text().split("\n")
# riccati_wb.cuh: slots < 20 are thread 0 (GEMM warp 0), slots >= 20 thread 256 (the helper warp): two separate time lines
for name, sel in [("GEMM warp 0" if not legacy else "thread 0", lambda sl: sl < 20), ("helper warp", lambda sl: sl >= 20)]:
    tot = sum(a[sl, 1] for sl in range(32) if sel(sl))
    if tot == 0:
        continue
    print(f"\n== K2 {'riccati_kernel' if legacy else 'riccati_bwd_kernel'} {name} (instance 0, cycles summed over all stages): {tot} cycles")
    for slot in range(32):
        if sel(slot) and a[slot, 1] > 0:
            line = int(a[slot, 0])
            ctx = " | ".join(t.strip()[:70] for t in src[max(0, line - 4):line - 1])
            print(f"  slot {slot:2d} {a[slot, 1]:9d} {100.0 * a[slot, 1] / tot:5.1f}%  L{line:<4d} {ctx}")
