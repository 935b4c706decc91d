"""Experiment: the same 256 instances as one handle on one stream vs. split over S handles on S CUDA streams driven by S host threads
(K2's latency-bound sweep of one part can overlap the throughput-bound K1 / K3 of another)."""
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from wb_humanoid_mpc_b200 import abi, model_loader  # noqa: E402
from wb_humanoid_mpc_b200.solver import B200SqpSolver, stack_instances  # noqa: E402

model = model_loader.load_packaged_model()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
insts = bench.build_batch(model, B, 0, 3.5, ["walk"])
st = abi.default_settings(model, sqp_iteration=1)


def run(parts, steps=10, warmup=3):
    solvers, streams = [], []
    per = B // parts
    for p in range(parts):
        s = B200SqpSolver(model, st)
        s.upload(stack_instances(insts[p * per:(p + 1) * per]))
        solvers.append(s)
        streams.append(torch.cuda.Stream())

    def worker(p, n):
        for _ in range(n):
            solvers[p].reset()
            solvers[p].solve(streams[p].cuda_stream)

    def go(n):
        th = [threading.Thread(target=worker, args=(p, n)) for p in range(parts)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize()

    go(warmup)
    t = time.perf_counter()
    go(steps)
    dt = time.perf_counter() - t
    x = np.concatenate([s.primal_solution()["x"] for s in solvers])
    return B * steps / dt, x


base, x1 = run(1)
print("1 stream : %.0f solves/s" % base)
for parts in (2, 4):
    v, x = run(parts)
    print("%d streams: %.0f solves/s (x%.3f), max |dx| vs 1 stream %.2e" % (parts, v, v / base, np.abs(x - x1).max()))
