"""Experiment (GPU box): bandwidth of b200sqp_upload_instances / b200sqp_download against a plain pinned torch copy of the same bytes
(found the row-granular cudaMemcpy2D slowdown that the contiguous-copy path for unpadded states fixes)."""
import sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import bench
from wb_humanoid_mpc_b200 import abi, model_loader
from wb_humanoid_mpc_b200.solver import B200SqpSolver, stack_instances
model = model_loader.load_packaged_model()
insts = bench.build_batch(model, 256, 0, 3.5, ["walk"])
batch = stack_instances(insts)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
pinned = {k: pin(v if v.dtype == np.uint8 else v.astype(np.float64)) for k, v in batch.items()}
s = B200SqpSolver(model, abi.default_settings(model, sqp_iteration=1))
s.upload(pinned)
t = time.perf_counter()
for _ in range(20): s.upload(pinned)
dt = (time.perf_counter() - t) / 20
nb = sum(v.nbytes for v in pinned.values())
print("upload %.3f ms  %.1f GB/s (%d MB)" % (dt * 1e3, nb / dt / 1e9, nb >> 20))
big = torch.empty(nb, dtype=torch.uint8).pin_memory(); dev = torch.empty(nb, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(20): dev.copy_(big, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 20
print("torch pinned H2D %.3f ms  %.1f GB/s" % (dt * 1e3, nb / dt / 1e9))
s.solve(); 
t = time.perf_counter()
for _ in range(20): s.primal_solution()
print("download (pageable out) %.3f ms" % ((time.perf_counter() - t) / 20 * 1e3))
