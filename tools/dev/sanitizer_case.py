"""compute-sanitizer case: one small whole-body solve with feedback gains and value function (memcheck / racecheck run on a GPU box:
compute-sanitizer --tool memcheck python tools/dev/sanitizer_case.py)."""
import numpy as np, sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from wb_humanoid_mpc_b200 import abi, model_loader
from wb_humanoid_mpc_b200.solver import B200SqpSolver
from test_gpu_wb import make_instances
model = model_loader.load_packaged_model()
insts = make_instances(model, np.random.default_rng(0), [("stance", 0.4, None), ("walk", 0.4, [0.4, 0.0, 0.7925, 0.0]), ("walk", 0.4, [0.1, 0.1, 0.7925, 0.2])])
# instances must share node count: build separately
for inst in insts[:2]:
    st = abi.default_settings(model, sqp_iteration=2, create_value_function=1, use_feedback_policy=1)
    s = B200SqpSolver(model, st)
    r = s.run([inst])
    print("ok", r["log"][0,:,8])
# a batch above the cluster threshold (18): riccati_bwd_kernel / riccati_fwd_kernel, the line-search ladder with several instances per CTA row
rng = np.random.default_rng(1)
batch = make_instances(model, rng, [("walk", 0.4, [0.3, 0.0, 0.7925, 0.1])] * 20)
st = abi.default_settings(model, sqp_iteration=2, create_value_function=1, use_feedback_policy=1)
s = B200SqpSolver(model, st)
r = s.run(batch)
print("ok batch", r["x"].shape, r["status"].tolist(), r["log"][:, 0, 8].tolist())
