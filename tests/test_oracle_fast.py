"""Pins the fast CPU baseline (oracle/fast/wb_fast.cu -- the TIMED arm of bench.py) on the checker oracle (oracle/wb_problem.hpp + sqp.hpp,
dense dual-number Jacobians): same iterate, gains and iteration log, so that the number bench.py reports as `cpu_baseline` is the time of
the same algorithm the parity tests check."""
import numpy as np
import pytest

import oracle_lib as orc
from wb_humanoid_mpc_b200 import abi, model_loader, references
from wb_humanoid_mpc_b200.solver import stack_instances


@pytest.fixture(scope="module")
def model():
    return model_loader.load_packaged_model()


def rel(a, b):
    return np.max(np.abs(a - b)) / max(1e-9, np.max(np.abs(b))) if b.size else 0.0


def make(model, rng, gait, horizon, cmd):
    x0 = np.array(model["x_init"], float)
    x0[2] = model["reference"]["defaultBaseHeight"]
    x0[0:2] += rng.uniform(-0.02, 0.02, 2)
    x0[3:6] += rng.uniform(-0.05, 0.05, 3)
    x0[6:29] += rng.uniform(-0.1, 0.1, 23)
    x0[29:] += rng.uniform(-0.2, 0.2, 29)
    return references.build_instance(model, x0, t0=0.0, horizon=horizon, gait=gait, cmd=cmd)


def oracle_solve(model, inst, st):
    wb = orc.WbOracle(model)
    wb.set_nodes(inst["contact_flags"], inst["swing_ref"], inst["impact_factor"], inst["arm_phase"], inst["x_ref"])
    return wb.sqp(inst["t_nodes"], inst["node_event"], inst["x0"], inst["x_init"], inst["u_init"], st)


@pytest.mark.parametrize("gait,cmd,iters,node_threads,seed", [("walk", [0.4, 0.0, 0.7925, 0.1], 1, 1, 11), ("stance", None, 3, 4, 12),
                                                               ("run", [0.3, 0.0, 0.7925, 0.0], 2, 2, 13)])
def test_fast_cpu_baseline_matches_checker_oracle(model, gait, cmd, iters, node_threads, seed):
    rng = np.random.default_rng(seed)
    inst = make(model, rng, gait, 0.5, cmd)
    st = abi.default_settings(model, sqp_iteration=iters, use_feedback_policy=1)
    ref = oracle_solve(model, inst, st)
    out = orc.fast_wb_sqp_batch(model, stack_instances([inst]), st, threads=1, node_threads=node_threads, want_gains=True)
    nit = len(ref["log"])
    assert out["n_iter"][0] == nit
    for it in range(nit):
        g, o = out["log"][0, it], ref["log"][it]
        assert g[8] == o[8] and int(g[9]) == int(o[9]) and int(g[13]) == int(o[13]), (it, g[8:14], o[8:14])
        for j in (0, 1, 2, 3, 4, 5, 6, 7, 10, 11):
            assert abs(g[j] - o[j]) <= 1e-7 * max(1.0, abs(o[j])), (it, j, g[j], o[j])
    assert rel(out["x"][0], ref["x"]) < 1e-7
    assert rel(out["u"][0], ref["u"]) < 1e-6
    assert rel(out["K"][0], ref["K"]) < 1e-5


def test_fast_cpu_baseline_batch_threads_agree(model):
    """instances spread over worker threads give the same answers as one after the other (no shared scratch)"""
    rng = np.random.default_rng(5)
    insts = [make(model, rng, "walk", 0.5, [0.3, 0.0, 0.7925, 0.0]) for _ in range(4)]
    st = abi.default_settings(model, sqp_iteration=1)
    a = orc.fast_wb_sqp_batch(model, stack_instances(insts), st, threads=1)
    b = orc.fast_wb_sqp_batch(model, stack_instances(insts), st, threads=4)
    assert np.array_equal(a["x"], b["x"]) and np.array_equal(a["u"], b["u"])
    assert a["stage_s"].min() > 0
