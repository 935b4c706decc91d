"""Experiment (GPU box): host-thread count of b200sqp::host::SqpSolver and the double-buffering pattern -- one solver object vs. two objects on
two host threads with setExclusiveSolve; prints ms per run() and the host-side stage split.  Result recorded in DESIGN.md section 5/6."""
import sys, threading, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import bench
from wb_humanoid_mpc_b200 import abi, host_lib, model_loader
model = model_loader.load_packaged_model()
insts = bench.build_batch(model, 256, 0, 3.5, ["walk"])
st = abi.default_settings(model, sqp_iteration=1)
hm = host_lib.HostModel()
x0s = np.array([i["x0"] for i in insts])
def mk(threads):
    hs = host_lib.HostSqpSolver(hm, st, 256, device=0, threads=threads)
    for b, i in enumerate(insts):
        hs.set_gait(b, i["gait"], 0.0, 10.5); hs.set_command(b, 0.0, i["x0"], i["cmd"], 3.5)
    return hs
def work(hs, n):
    for _ in range(n):
        hs.reset(); hs.run(0.0, x0s, 3.5)
for threads in (8, 16, 32):
    a, b = mk(threads), mk(threads)
    a.set_exclusive_solve(True); b.set_exclusive_solve(True)
    work(a, 2); work(b, 2)
    t = time.perf_counter(); work(a, 6); ser = (time.perf_counter() - t) / 6
    th = [threading.Thread(target=work, args=(s, 3)) for s in (a, b)]
    t = time.perf_counter(); [x.start() for x in th]; [x.join() for x in th]; pip = (time.perf_counter() - t) / 6
    print("threads %3d: serial %.1f ms/run (%.0f solves/s)  2 solvers %.1f ms/run (%.0f solves/s)  stages %s" % (threads, ser * 1e3, 256 / ser, pip * 1e3, 256 / pip, [round(v, 2) for v in a.benchmarks()[4:10]]))
    a.close(); b.close()
