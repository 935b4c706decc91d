"""Device-side instance builder (b200sqp_build_instances) against the host restatement of SolverBase::preRun
(wb_humanoid_mpc_b200/references.py, itself pinned on the oracle and the C++ host layer): every per-node array the GPU builds from (x0, gait,
gait start, velocity command) equals what the host builds and uploads -- integer arrays exactly, floating-point arrays to round-off (the
two compilers contract multiply-adds differently) -- for cold starts of several gaits, for the reference's warm start from the previous
solution left on the device, and the solve that follows gives the same answer as the upload path."""
import numpy as np
import pytest

from wb_humanoid_mpc_b200 import abi, model_loader, references
from wb_humanoid_mpc_b200.lib import B200SqpError

pytestmark = pytest.mark.gpu
FLOAT_KEYS = ["x0", "x_init", "u_init", "t_nodes", "swing_ref", "impact_factor", "arm_phase", "x_ref"]
INT_KEYS = ["node_event", "contact_flags"]


@pytest.fixture(scope="module")
def model():
    return model_loader.load_packaged_model()


def make_inputs(model, rng, B):
    x0s, cmds = [], []
    for _ in range(B):
        x0 = np.array(model["x_init"], float)
        x0[2] = model["reference"]["defaultBaseHeight"]
        x0[0:2] += rng.uniform(-0.02, 0.02, 2)
        x0[3:6] += rng.uniform(-0.05, 0.05, 3)
        x0[6:29] += rng.uniform(-0.1, 0.1, 23)
        x0[29:] += rng.uniform(-0.2, 0.2, 29)
        x0s.append(x0)
        cmds.append([rng.uniform(-0.5, 1.0), rng.uniform(-0.3, 0.3), model["reference"]["defaultBaseHeight"], rng.uniform(-0.5, 0.5)])
    return np.array(x0s), np.array(cmds)


def compare(dev, host_insts):
    from wb_humanoid_mpc_b200.solver import stack_instances

    ref = stack_instances(host_insts)
    for k in INT_KEYS:
        assert np.array_equal(dev[k], ref[k]), k
    for k in FLOAT_KEYS:
        err = np.max(np.abs(dev[k] - ref[k]))
        assert err <= 1e-12 * max(1.0, np.max(np.abs(ref[k]))), (k, err)


@pytest.mark.parametrize("gait,start,horizon", [("walk", 0.0, 3.5), ("stance", 0.5, 1.1), ("run", -0.23, 1.1), ("jump", 0.0, 0.8), ("slow_walk", -1.1, 3.5),
                                                ("trot", 0.3, 1.1)])
def test_cold_start_arrays_match_the_host_builder(model, gait, start, horizon):
    from wb_humanoid_mpc_b200.solver import B200SqpSolver

    rng = np.random.default_rng(abs(int(start * 100)) + len(gait))
    B = 5
    x0s, cmds = make_inputs(model, rng, B)
    solver = B200SqpSolver(model, abi.default_settings(model, sqp_iteration=1))
    n = solver.build_instances(0.0, horizon, x0s, [gait] * B, [start] * B, cmds)
    dev = solver.download_instances()
    host = [references.build_instance(model, x0s[b], t0=0.0, horizon=horizon, gait=gait, gait_start=(None if gait == "stance" else start), cmd=list(cmds[b]))
            for b in range(B)]
    assert n == len(host[0]["t_nodes"])
    compare(dev, host)


def test_solve_after_device_build_equals_the_upload_path(model):
    from wb_humanoid_mpc_b200.solver import B200SqpSolver

    rng = np.random.default_rng(7)
    B = 6
    x0s, cmds = make_inputs(model, rng, B)
    st = abi.default_settings(model, sqp_iteration=2)
    a = B200SqpSolver(model, st)
    a.build_instances(0.0, 1.1, x0s, ["walk"] * B, [0.0] * B, cmds)
    a.solve()
    ra = a.primal_solution()
    host = [references.build_instance(model, x0s[b], t0=0.0, horizon=1.1, gait="walk", cmd=list(cmds[b])) for b in range(B)]
    rb = B200SqpSolver(model, st).run(host)
    assert np.array_equal(ra["log"][:, :, 8], rb["log"][:, :, 8])
    # the two input sets agree to 1e-16 relative; that is enough to resolve an exact tie of the complete-pivoting LU differently at some node, and
    # the two (equally valid) bases of the projected QP then give answers that differ by its conditioning: 1e-5 relative
    assert np.max(np.abs(ra["x"] - rb["x"])) < 1e-5 * np.max(np.abs(rb["x"])) and np.max(np.abs(ra["u"] - rb["u"])) < 1e-5 * np.max(np.abs(rb["u"]))
    for j in (0, 1, 2, 3, 4, 5, 6, 7):
        assert np.allclose(ra["log"][:, :, j], rb["log"][:, :, j], rtol=1e-7, atol=1e-9)


def test_receding_horizon_on_the_device(model):
    """three MPC cycles without a host copy of x / u: build (cold) -> solve -> build (warm, shifted on the device) -> solve -> ...; after every
    warm build the arrays equal the host's warm start from the same previous solution (Initialization.cpp:35-79), and the new event nodes that
    enter the horizon change the node count"""
    from wb_humanoid_mpc_b200.solver import B200SqpSolver

    rng = np.random.default_rng(9)
    B, horizon, dt_mpc = 4, 1.1, 0.1
    x0s, cmds = make_inputs(model, rng, B)
    solver = B200SqpSolver(model, abi.default_settings(model, sqp_iteration=1))
    solver.build_instances(0.0, horizon, x0s, ["walk"] * B, [0.0] * B, cmds)
    counts = set()
    for cycle in range(1, 4):
        solver.solve()
        prev_inputs = solver.download_instances()
        sol = solver.primal_solution()
        t1 = cycle * dt_mpc
        prevs = [references.to_primal_solution(prev_inputs["t_nodes"][b], prev_inputs["node_event"][b], sol["x"][b], sol["u"][b]) for b in range(B)]
        x1 = np.array([references.linear_interpolate(t1, p["t"], p["x"]) for p in prevs])
        n = solver.build_instances(t1, horizon, x1, ["walk"] * B, [0.0] * B, cmds, warm=True)
        counts.add(n)
        dev = solver.download_instances()
        host = [references.build_instance(model, x1[b], t0=t1, horizon=horizon, gait="walk", gait_start=0.0, cmd=list(cmds[b]), previous=prevs[b]) for b in range(B)]
        assert n == len(host[0]["t_nodes"])
        compare(dev, host)
        assert not np.allclose(dev["u_init"][:, 0], dev["u_init"][:, -1])   # the overlap is interpolated, the tail is the initializer
    assert len(counts) >= 1


def test_closed_loop_without_state_transfer(model):
    """x0 = NULL with warm = 1: the measured state of the next cycle is the plan interpolated at the new initial time on the device; 12 cycles of
    a walking batch stay feasible and the constraint violation of the real-time iteration falls"""
    from wb_humanoid_mpc_b200.solver import B200SqpSolver

    rng = np.random.default_rng(11)
    B, horizon = 8, 1.1
    x0s, cmds = make_inputs(model, rng, B)
    solver = B200SqpSolver(model, abi.default_settings(model, sqp_iteration=1))
    solver.build_instances(0.0, horizon, x0s, ["walk"] * B, [0.0] * B, cmds)
    solver.solve()
    first = solver.iterations_log()[:, 0]
    prev = solver.primal_solution()
    prev_t = solver.download_instances()["t_nodes"]
    for c in range(1, 13):
        t1 = c / 60.0
        solver.build_instances(t1, horizon, None, ["walk"] * B, [0.0] * B, cmds, warm=True)
        inp = solver.download_instances()
        # the state the device took as measurement is the previous plan at t1
        want = np.array([references.linear_interpolate(t1, prev_t[b], prev["x"][b]) for b in range(B)])
        assert np.max(np.abs(inp["x0"] - want)) < 1e-12
        solver.solve()
        prev = solver.primal_solution()
        prev_t = inp["t_nodes"]
        assert not prev["status"].any()
    last = solver.iterations_log()[:, 0]
    g0, g1 = np.sqrt(first[:, 2] + first[:, 3]), np.sqrt(last[:, 2] + last[:, 3])
    assert np.median(g1) < 0.2 * np.median(g0)


def test_mixed_node_counts_are_refused(model):
    from wb_humanoid_mpc_b200.solver import B200SqpSolver

    rng = np.random.default_rng(1)
    x0s, cmds = make_inputs(model, rng, 2)
    solver = B200SqpSolver(model, abi.default_settings(model, sqp_iteration=1))
    with pytest.raises(B200SqpError) as e:
        solver.build_instances(0.0, 3.5, x0s, ["walk", "trot"], [0.0, 0.0], cmds)
    assert e.value.code == -1 and "shooting nodes" in str(e.value)
    with pytest.raises(B200SqpError) as e2:
        solver.build_instances(0.0, 3.5, x0s, ["walk", "walk"], [0.0, 0.7], cmds)
    assert e2.value.code == -1
