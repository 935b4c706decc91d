"""b200sqp::host::SqpSolver (C++ mirror of ocs2::SqpSolver over the C ABI) against the Python harness: same instances -> same primal
(the two instance builders agree to 1e-13, not bitwise, so the solutions agree to solver conditioning: 1e-9 on x)
solution, iteration log and warm-started second solve (grouping by node count: the last test of this file, where the gaits really differ in
their node counts)."""
import numpy as np
import pytest

from wb_humanoid_mpc_b200 import abi, host_lib, model_loader, references

pytestmark = pytest.mark.gpu


def test_host_solver_matches_python_harness_cold_and_warm():
    from wb_humanoid_mpc_b200.solver import LOG_FIELDS, B200SqpSolver

    model = model_loader.load_packaged_model()
    hm = host_lib.HostModel()
    rng = np.random.default_rng(3)
    B, horizon = 5, 1.1
    gaits = ["walk", "stance", "walk", "trot", "stance"]
    cmds = [[rng.uniform(-0.3, 0.8), rng.uniform(-0.2, 0.2), model["reference"]["defaultBaseHeight"], rng.uniform(-0.3, 0.3)] for _ in range(B)]
    x0s = []
    for _ in range(B):
        x0 = np.array(model["x_init"], float)
        x0[2] = model["reference"]["defaultBaseHeight"]
        x0[3:6] += rng.uniform(-0.05, 0.05, 3)
        x0[6:29] += rng.uniform(-0.05, 0.05, 23)
        x0s.append(x0)
    st = abi.default_settings(model, sqp_iteration=2)
    host = host_lib.HostSqpSolver(hm, st, B)
    for b in range(B):
        host.set_gait(b, gaits[b], 0.0, 3 * horizon)
        host.set_command(b, 0.0, x0s[b], cmds[b], horizon)
    host.run(0.0, np.array(x0s), horizon)

    prev = []
    for b in range(B):
        inst = references.build_instance(model, x0s[b], t0=0.0, horizon=horizon, gait=gaits[b], cmd=cmds[b])
        py = B200SqpSolver(model, st)
        r = py.run([inst])
        p = host.primal_solution(b)
        assert np.array_equal(p["t"], inst["t_nodes"])
        dx = float(np.abs(p["x"] - r["x"][0]).max())
        print("instance", b, gaits[b], "max |dx| host vs python", dx)
        assert dx < 1e-6, (b, dx)
        ref_u = references.to_primal_solution(inst["t_nodes"], inst["node_event"], r["x"][0], r["u"][0])["u"]
        assert np.allclose(p["u"], ref_u, rtol=0, atol=1e-4 * max(1.0, np.abs(ref_u).max())), float(np.abs(p["u"] - ref_u).max())
        lg = host.iterations_log(b)
        assert lg.shape[0] == r["n_iter"][0]
        assert np.allclose(lg[:, 6], r["log"][0, : lg.shape[0], LOG_FIELDS.index("step_size")])
        assert np.allclose(lg[:, 3], r["log"][0, : lg.shape[0], LOG_FIELDS.index("merit")], rtol=1e-7)
        prev.append((py, references.to_primal_solution(inst["t_nodes"], inst["node_event"], r["x"][0], r["u"][0], inst["mode_schedule"])))
    assert host.benchmarks()[0] > 0.0
    # sqp::Logger CSV (host/SqpLogging.hpp): the reference's 19 columns, one line per instance and iteration
    csv = host.write_log(0.0).strip().split("\n")
    assert csv[0].split(", ")[:3] == ["problemNumber", "time", "iteration"] and len(csv[0].split(", ")) == 19
    assert len(csv) - 1 == sum(host.iterations_log(b).shape[0] for b in range(B))
    row = csv[1].split(", ")
    assert len(row) == 19 and int(row[0]) == 0 and float(row[10]) == host.iterations_log(0)[0, 6] and row[11] in ("Constraint", "Dual", "Cost", "Zero")
    assert row[18] in ("Not Converged", "Maximum number of iterations reached")

    # receding horizon: second solve 3 nodes later, warm-started from the previous primal solution on both sides
    t1 = 3 * model["sqp"]["dt"]
    x1s = [prev[b][1]["x"][3] for b in range(B)]
    for b in range(B):
        host.set_command(b, t1, x1s[b], cmds[b], horizon)
    host.run(t1, np.array(x1s), t1 + horizon)
    for b in range(B):
        inst = references.build_instance(model, x1s[b], t0=t1, horizon=horizon, gait=gaits[b], gait_start=0.0, cmd=cmds[b], previous=prev[b][1])
        r = prev[b][0].run([inst])
        p = host.primal_solution(b)
        dx = float(np.abs(p["x"] - r["x"][0]).max())
        print("warm instance", b, "max |dx|", dx)
        assert dx < 1e-6, (b, dx)
    host.close()
    hm.close()


@pytest.mark.parametrize("spread", [True, False])
def test_receding_horizon_converges(spread):
    """closed loop with perfect tracking (tools/bench_receding.py in small): the real-time-iteration scheme warm-started from the previous
    primal solution drives the dynamics / constraint violations of the plan down over the MPC cycles and reaches full steps.
    spread=True is the reference behaviour: trajectorySpread moves the time stamp of the sample after every event (its convention clash with
    the SQP's primal solution), which distorts the warm start next to events and slows the decrease; spread=False keeps the stamps."""
    model = model_loader.load_packaged_model()
    hm = host_lib.HostModel()
    st = abi.default_settings(model, sqp_iteration=1)
    B, T, dt = 4, 1.1, model["sqp"]["dt"]
    host = host_lib.HostSqpSolver(hm, st, B)
    host.set_trajectory_spread(spread)
    rng = np.random.default_rng(8)
    x = np.array([np.array(model["x_init"], float) for _ in range(B)])
    x[:, 2] = model["reference"]["defaultBaseHeight"]
    x[:, 6:29] += rng.uniform(-0.05, 0.05, (B, 23))
    cmds = [[0.4, 0.0, model["reference"]["defaultBaseHeight"], 0.1]] * B
    for b in range(B):
        host.set_gait(b, "walk", 0.0, 20 * dt + 3 * T)
    t, viol, steps = 0.0, [], []
    for c in range(10):
        for b in range(B):
            host.set_command(b, t, x[b], cmds[b], T)
        host.run(t, x, t + T)
        logs = np.array([host.iterations_log(b)[0] for b in range(B)])
        assert np.all(np.isfinite(logs))
        viol.append(float((logs[:, 4] + logs[:, 5]).mean()))
        steps.append(float(logs[:, 6].mean()))
        x = np.array([host.primal_solution(b)["x"][1] for b in range(B)])
        t += dt
    if spread:
        assert min(viol) < 0.25 * viol[0] and viol[-1] < 0.5 * viol[0], viol
    else:
        assert viol[-1] < 0.2 * viol[0] and all(b < a * 1.5 for a, b in zip(viol, viol[1:])), viol
        assert steps[-1] >= steps[0] and steps[-1] > 0.7, steps
    host.close()
    hm.close()


def test_host_solver_centroidal_matches_python_harness():
    """the C++ host layer on the centroidal model file: b200sqp_cen_create behind b200sqp::host::SqpSolver, cold and warm-started solve"""
    from wb_humanoid_mpc_b200 import centroidal
    from wb_humanoid_mpc_b200.solver import B200SqpSolver

    model = model_loader.load_packaged_model("g1_centroidal")
    hm = host_lib.HostModel(host_lib.CEN_MODEL_TXT)
    rng = np.random.default_rng(5)
    B, horizon = 3, 0.6
    gaits = ["walk", "stance", "walk"]
    cmds = [[rng.uniform(-0.3, 0.8), rng.uniform(-0.2, 0.2), model["reference"]["defaultBaseHeight"], rng.uniform(-0.3, 0.3)] for _ in range(B)]
    x0s = []
    for _ in range(B):
        x0 = np.array(model["x_init"], float)
        x0[:6] = rng.uniform(-0.05, 0.05, 6)
        x0[9:12] += rng.uniform(-0.05, 0.05, 3)
        x0[12:] += rng.uniform(-0.05, 0.05, 23)
        x0s.append(x0)
    st = abi.default_settings(model, sqp_iteration=2)
    host = host_lib.HostSqpSolver(hm, st, B)
    bvs = [hm.base_velocity(x0) for x0 in x0s]
    for b in range(B):
        assert np.allclose(bvs[b], centroidal.base_velocity(model, x0s[b]), atol=1e-14)
        host.set_gait(b, gaits[b], 0.0, 3 * horizon)
        host.set_command(b, 0.0, x0s[b], cmds[b], horizon, base_vel=bvs[b])
    host.run(0.0, np.array(x0s), horizon)
    prev = []
    for b in range(B):
        inst = references.build_instance(model, x0s[b], t0=0.0, horizon=horizon, gait=gaits[b], cmd=cmds[b], base_vel=bvs[b])
        py = B200SqpSolver(model, st)
        r = py.run([inst])
        p = host.primal_solution(b)
        assert p["x"].shape[1] == 35 and np.array_equal(p["t"], inst["t_nodes"])
        dx = float(np.abs(p["x"] - r["x"][0]).max())
        print("centroidal instance", b, gaits[b], "max |dx| host vs python", dx)
        assert dx < 1e-6, (b, dx)
        prev.append((py, references.to_primal_solution(inst["t_nodes"], inst["node_event"], r["x"][0], r["u"][0], inst["mode_schedule"])))
    t1 = 3 * model["sqp"]["dt"]
    x1s = [prev[b][1]["x"][3] for b in range(B)]
    bv1 = [hm.base_velocity(x) for x in x1s]
    for b in range(B):
        host.set_command(b, t1, x1s[b], cmds[b], horizon, base_vel=bv1[b])
    host.run(t1, np.array(x1s), t1 + horizon)
    for b in range(B):
        inst = references.build_instance(model, x1s[b], t0=t1, horizon=horizon, gait=gaits[b], gait_start=0.0, cmd=cmds[b], previous=prev[b][1],
                                         base_vel=bv1[b])
        r = prev[b][0].run([inst])
        dx = float(np.abs(host.primal_solution(b)["x"] - r["x"][0]).max())
        print("centroidal warm instance", b, "max |dx|", dx)
        assert dx < 1e-6, (b, dx)
    host.close()
    hm.close()


def test_double_buffered_solvers_and_concurrent_groups_reproduce_the_serial_result():
    """two SqpSolver objects on two host threads with setExclusiveSolve (the double-buffering pattern of bench.py) and a mixed batch whose
    node-count groups are solved concurrently: every instance must come out exactly as from a single solver run on its own"""
    import threading

    model = model_loader.load_packaged_model()
    hm = host_lib.HostModel()
    rng = np.random.default_rng(21)
    B, horizon = 6, 1.5
    gaits = ["walk", "stance", "slow_walk", "walk", "trot", "stance"]   # node counts 50 / 48 / 47 at this horizon -> three groups per solver
    x0s = []
    for _ in range(B):
        x0 = np.array(model["x_init"], float)
        x0[2] = model["reference"]["defaultBaseHeight"]
        x0[6:29] += rng.uniform(-0.05, 0.05, 23)
        x0s.append(x0)
    x0s = np.array(x0s)
    st = abi.default_settings(model, sqp_iteration=2)

    def make():
        s = host_lib.HostSqpSolver(hm, st, B)
        for b in range(B):
            s.set_gait(b, gaits[b], 0.0, 3 * horizon)
            s.set_command(b, 0.0, x0s[b], [0.3, 0.0, model["reference"]["defaultBaseHeight"], 0.1], horizon)
        return s

    ref = make()
    ref.set_exclusive_solve(True)   # exclusive mode solves the groups one after the other: the serial reference
    ref.run(0.0, x0s, horizon)
    want = [ref.primal_solution(b)["x"].copy() for b in range(B)]
    assert len({w.shape[0] for w in want}) > 1   # the batch really has several node counts

    conc = make()                   # default mode: groups concurrently, one host thread and CUDA stream per group
    conc.run(0.0, x0s, horizon)
    for b in range(B):
        assert np.array_equal(conc.primal_solution(b)["x"], want[b]), b

    pair = [make(), make()]
    for s in pair:
        s.set_exclusive_solve(True)

    def work(s):
        for _ in range(3):
            s.reset()
            s.run(0.0, x0s, horizon)

    th = [threading.Thread(target=work, args=(s,)) for s in pair]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for s in pair:
        for b in range(B):
            assert np.array_equal(s.primal_solution(b)["x"], want[b]), b
        s.close()
    ref.close()
    conc.close()
    hm.close()
