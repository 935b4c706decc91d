"""torchrun worker of tests/test_gpu_global_step.py::test_two_rank_nccl: a global batch sharded over 2 ranks by batch index; the global-step
mode must pick, on both ranks, the step a single process picks for the whole batch."""
import os

import numpy as np
import torch
import torch.distributed as dist

from wb_humanoid_mpc_b200 import abi, model_loader, parallel
from wb_humanoid_mpc_b200.solver import LOG_FIELDS, B200SqpSolver
from test_gpu_wb import make_instances

def lg(res, name):
    return res["log"][:, :, LOG_FIELDS.index(name)]


rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
model = model_loader.load_packaged_model()
rng = np.random.default_rng(23)
specs = [("walk", 1.1, [rng.uniform(-0.5, 1.0), rng.uniform(-0.3, 0.3), 0.7925, rng.uniform(-0.5, 0.5)]) for _ in range(8)]
insts = make_instances(model, rng, specs)          # the same global batch on every rank
lo, hi = parallel.shard_bounds(len(insts), rank, world)
st = abi.default_settings(model, sqp_iteration=2, global_step=1)
sharded = B200SqpSolver(model, st, device=local)
sharded.enable_global_step()
r = sharded.run(insts[lo:hi])
whole = B200SqpSolver(model, st, device=local)     # reference: one process owning the whole batch, local decision
w = whole.run(insts)
assert np.array_equal(lg(r, "step_size"), lg(w, "step_size")[lo:hi]), (lg(r, "step_size"), lg(w, "step_size")[lo:hi])
assert np.allclose(r["x"], w["x"][lo:hi], rtol=0, atol=1e-12)
steps = torch.tensor(lg(r, "step_size")[0], device="cuda")
allsteps = [torch.empty_like(steps) for _ in range(world)]
dist.all_gather(allsteps, steps)
assert all(torch.equal(allsteps[0], s) for s in allsteps)
dist.barrier()
if rank == 0:
    print("GLOBAL_STEP_OK", lg(r, "step_size")[0].tolist())
dist.destroy_process_group()
