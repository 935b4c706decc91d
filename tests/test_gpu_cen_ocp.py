"""Centroidal SQP parity (GPU, through the C ABI) against the CPU oracle: BASELINE configs[0] (N = 20) and configs[1] (G1 full centroidal
dynamics, N = 100, batch 1, fp64 correctness).

Tolerances (fp64): LQ blocks 1e-9 relative; SQP outputs (accepted step, performance indices, primal trajectory, remapped gains, cost-to-go)
1e-7 relative on the first iterations, as for the whole-body path (tests/test_gpu_wb.py)."""
import numpy as np
import pytest

import oracle_lib as orc
from wb_humanoid_mpc_b200 import abi, centroidal, model_loader, references

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model():
    return model_loader.load_packaged_model("g1_centroidal")


def rel(a, b):
    return np.max(np.abs(a - b)) / max(1e-9, np.max(np.abs(b))) if b.size else 0.0


def make_instance(model, rng, gait, horizon, cmd, momentum=False):
    x0 = np.array(model["x_init"], float)
    x0[6:8] += rng.uniform(-0.02, 0.02, 2)
    x0[9:12] += rng.uniform(-0.05, 0.05, 3)
    x0[12:] += rng.uniform(-0.1, 0.1, model["nj"])
    base_vel = None
    if momentum:
        x0[:6] += rng.uniform(-0.1, 0.1, 6)
        base_vel = centroidal.base_velocity(model, x0)
    return references.build_instance(model, x0, t0=0.0, horizon=horizon, gait=gait, cmd=cmd, base_vel=base_vel)


def oracle_solve(model, inst, settings, keep_raw=False):
    o = orc.CenOracle(model)
    o.set_nodes(inst["contact_flags"], inst["swing_ref"], inst["impact_factor"], inst["arm_phase"], inst["x_ref"])
    res = o.sqp(inst["t_nodes"], inst["node_event"], inst["x0"], inst["x_init"], inst["u_init"], settings, keep_raw=keep_raw)
    if keep_raw:
        res["raw"] = o.last_raw_blocks(len(inst["t_nodes"]))
    res["oracle"] = o
    return res


def test_base_velocity_helper_matches_oracle(model):
    rng = np.random.default_rng(0)
    x0 = np.array(model["x_init"], float)
    x0[:6] = rng.uniform(-0.2, 0.2, 6)
    x0[9:12] += rng.uniform(-0.1, 0.1, 3)
    bv = centroidal.base_velocity(model, x0)
    o = orc.CenOracle(model)
    xd = o.flow_map(x0, np.zeros(model["nu"]))
    assert np.allclose(bv, xd[6:12] / sum(model["mass"]), atol=1e-12)


def test_lq_blocks_match_oracle_config0(model):
    """configs[0] shape: N = 20 (dt 0.02, horizon 0.4 s); every raw stage block against the oracle"""
    from wb_humanoid_mpc_b200.solver import B200SqpSolver

    rng = np.random.default_rng(1)
    insts = [make_instance(model, rng, "walk", 0.4, [0.3, 0.1, 0.7925, 0.2], momentum=True), make_instance(model, rng, "walk", 0.4, None)]
    assert len(insts[0]["t_nodes"]) == len(insts[1]["t_nodes"]) == 21
    st = abi.default_settings(model, sqp_iteration=1)
    solver = B200SqpSolver(model, st, capture_raw_blocks=True)
    solver.run(insts)
    raw = solver.raw_stage_blocks()
    for b, inst in enumerate(insts):
        ref = oracle_solve(model, inst, st, keep_raw=True)
        for k in range(len(inst["t_nodes"]) - 1):
            g = orc.unpack_raw_blocks(raw[b, k], 35, 35)
            o = ref["raw"][k]
            assert g["nc"] == o["nc"], (b, k)
            for key in ["A", "B", "b", "Q", "S", "R", "q", "r", "C", "D", "e"]:
                assert rel(g[key], o[key]) < 1e-9, (b, k, key, rel(g[key], o[key]))


@pytest.mark.parametrize("gait,iters", [("stance", 1), ("walk", 3)])
def test_sqp_matches_oracle(model, gait, iters):
    from wb_humanoid_mpc_b200.solver import B200SqpSolver

    rng = np.random.default_rng(2)
    inst = make_instance(model, rng, gait, 0.6, [0.3, 0.0, 0.7925, 0.1])
    st = abi.default_settings(model, sqp_iteration=iters, use_feedback_policy=1, create_value_function=1)
    solver = B200SqpSolver(model, st)
    x_lin = inst["x_init"]
    sol = solver.run([inst])
    assert not sol["status"].any()
    ref = oracle_solve(model, inst, st)
    nit = len(ref["log"])
    assert sol["n_iter"][0] == nit
    for it in range(nit):
        g, o = sol["log"][0, it], ref["log"][it]
        assert g[8] == o[8], ("step size", it, g[8], o[8])
        assert int(g[9]) == int(o[9]) and int(g[13]) == int(o[13])
        for j in (0, 1, 2, 3, 4, 5, 6, 7, 10, 11):
            assert abs(g[j] - o[j]) <= 1e-6 * max(1.0, abs(o[j])), (it, j, g[j], o[j])
    assert rel(sol["x"][0], ref["x"]) < 1e-6
    assert rel(sol["u"][0], ref["u"]) < 1e-6
    assert rel(sol["K"][0], ref["K"]) < 1e-5
    if iters == 1:
        P, p = solver.value_function()
        Po, po = ref["oracle"].last_value_function(x_lin)
        assert rel(P[0], Po) < 1e-6 and rel(p[0], po) < 1e-6


def test_config1_full_centroidal_n100(model):
    """configs[1]: G1 full centroidal dynamics, N = 100 (dt 0.02, horizon 2.0 s), batch 1: one SQP iteration against the oracle, then the
    size-independent property that the accepted step satisfies the linearised dynamics and projected constraints of every stage."""
    from wb_humanoid_mpc_b200.solver import B200SqpSolver

    rng = np.random.default_rng(3)
    inst = make_instance(model, rng, "walk", 2.0, [0.4, 0.0, 0.7925, 0.0])
    n = len(inst["t_nodes"])
    assert n >= 101
    st = abi.default_settings(model, sqp_iteration=1)
    solver = B200SqpSolver(model, st, capture_raw_blocks=True)
    sol = solver.run([inst])
    raw = solver.raw_stage_blocks()
    assert not sol["status"].any()
    ref = oracle_solve(model, inst, st)
    g, o = sol["log"][0, 0], ref["log"][0]
    assert g[8] == o[8] and int(g[9]) == int(o[9])
    for j in (0, 1, 2, 3, 4, 5, 6, 7, 10, 11):
        assert abs(g[j] - o[j]) <= 1e-6 * max(1.0, abs(o[j])), (j, g[j], o[j])
    assert rel(sol["x"][0], ref["x"]) < 1e-6 and rel(sol["u"][0], ref["u"]) < 1e-6
    alpha = g[8]
    assert alpha > 0
    dx = (sol["x"][0] - inst["x_init"]) / alpha
    du = (sol["u"][0] - inst["u_init"]) / alpha
    assert np.allclose(dx[0], inst["x0"] - inst["x_init"][0], atol=1e-9)
    for k in range(n - 1):
        blk = orc.unpack_raw_blocks(raw[0, k], 35, 35)
        if inst["node_event"][k] == 1:
            assert np.max(np.abs(dx[k] + blk["b"] - dx[k + 1])) < 1e-8
            continue
        assert np.max(np.abs(blk["A"] @ dx[k] + blk["B"] @ du[k] + blk["b"] - dx[k + 1])) < 1e-7
        assert np.max(np.abs(blk["C"] @ dx[k] + blk["D"] @ du[k] + blk["e"])) < 1e-6
    ms = solver.benchmarks()
    assert ms[0] > 0 and ms[1] > 0 and ms[2] > 0


def test_batch_of_instances_is_independent(model):
    """the batch dimension: each instance of a batch reproduces its single-instance solve bit for bit"""
    from wb_humanoid_mpc_b200.solver import B200SqpSolver

    rng = np.random.default_rng(4)
    insts = [make_instance(model, rng, "walk", 0.4, [0.2 * i, 0.0, 0.7925, 0.0]) for i in range(3)]
    st = abi.default_settings(model, sqp_iteration=2)
    sol = B200SqpSolver(model, st).run(insts)
    for b, inst in enumerate(insts):
        one = B200SqpSolver(model, st).run([inst])
        assert np.array_equal(one["x"][0], sol["x"][b]) and np.array_equal(one["u"][0], sol["u"][b])


def test_srbd_model_type_matches_oracle(model):
    """centroidalModelType 1 (SingleRigidBodyDynamics): LQ blocks and two SQP iterations against the oracle"""
    from wb_humanoid_mpc_b200.solver import B200SqpSolver

    m1 = dict(model)
    m1["centroidalModelType"] = 1
    rng = np.random.default_rng(6)
    inst = make_instance(m1, rng, "walk", 0.4, [0.3, 0.1, 0.7925, 0.2], momentum=True)
    st = abi.default_settings(m1, sqp_iteration=2)
    solver = B200SqpSolver(m1, st, capture_raw_blocks=True)
    sol = solver.run([inst])
    ref = oracle_solve(m1, inst, st, keep_raw=True)
    assert sol["n_iter"][0] == len(ref["log"])
    for it in range(len(ref["log"])):
        g, o = sol["log"][0, it], ref["log"][it]
        assert g[8] == o[8]
        for j in (0, 1, 2, 3, 4, 5, 6, 7, 10, 11):
            assert abs(g[j] - o[j]) <= 1e-6 * max(1.0, abs(o[j])), (it, j, g[j], o[j])
    assert rel(sol["x"][0], ref["x"]) < 1e-6 and rel(sol["u"][0], ref["u"]) < 1e-6
    raw = solver.raw_stage_blocks()   # blocks of the LAST iteration on both sides
    for k in range(len(inst["t_nodes"]) - 1):
        g, o = orc.unpack_raw_blocks(raw[0, k], 35, 35), ref["raw"][k]
        for key in ["A", "B", "b", "C", "D", "e", "Q", "R"]:
            assert rel(g[key], o[key]) < 1e-6, (k, key)
    # and the model type matters: the full model gives different dynamics blocks
    full = B200SqpSolver(model, st, capture_raw_blocks=True)
    full.run([inst])
    assert rel(orc.unpack_raw_blocks(full.raw_stage_blocks()[0, 0], 35, 35)["B"], ref["raw"][0]["B"]) > 1e-4
