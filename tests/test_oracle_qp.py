"""Pins the CPU oracle's QP/SQP layer against the reference's own known-answer recipes (SURVEY.md §8c i-ix).

The reference stores no numeric golden files; each of its solver tests builds a seeded random LQ problem and checks a
property at 1e-9 ... 1e-12.  The recipes are re-run here on the oracle with numpy-built inputs.  The "textbook" Riccati
recursion below is written independently in numpy exactly as hpipm_catkin/test/testHpipmInterface.cpp:281-304 does.
"""
import numpy as np
import pytest

import oracle_lib as orc

TOL = 1e-9


def rand_cost(rng, n, m):
    """getRandomCost (ocs2_oc/test/include/ocs2_oc/test/testProblemsGeneration.h:45-59)"""
    M = rng.uniform(-1, 1, (n + m, n + m))
    M = M.T @ M
    return dict(Q=M[:n, :n], S=M[n:, :n], R=M[n:, n:], q=rng.uniform(-1, 1, n), r=rng.uniform(-1, 1, m))


def rand_dyn(rng, n, m):
    """getRandomDynamics (:71-79)"""
    return dict(A=rng.uniform(-1, 1, (n, n)), B=rng.uniform(-1, 1, (n, m)), b=rng.uniform(-1, 1, n))


def stack(dyn, cost, numax):
    N, nx = len(dyn), dyn[0]["A"].shape[0]
    A = np.stack([d["A"] for d in dyn])
    B = np.zeros((N, nx, numax))
    S = np.zeros((N, numax, nx))
    R = np.zeros((N, numax, numax))
    r = np.zeros((N, numax))
    nu = np.zeros(N, dtype=np.int32)
    for k in range(N):
        m = dyn[k]["B"].shape[1]
        nu[k] = m
        B[k, :, :m] = dyn[k]["B"]
        S[k, :m] = cost[k]["S"]
        R[k, :m, :m] = cost[k]["R"]
        r[k, :m] = cost[k]["r"]
    b = np.stack([d["b"] for d in dyn])
    Q = np.stack([c["Q"] for c in cost])
    q = np.stack([c["q"] for c in cost])
    return A, B, b, Q, S, R, q, r, nu


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_retrieve_riccati(seed):
    """testHpipmInterface.cpp:258-340 retrieveRiccati: P, p, K, k vs the textbook recursion; u = K x + k."""
    rng = np.random.default_rng(seed)
    nx, nu, N = 3, 2, 5
    x0 = rng.uniform(-1, 1, nx)
    dyn = [rand_dyn(rng, nx, nu) for _ in range(N)]
    cost = [rand_cost(rng, nx, nu) for _ in range(N)] + [rand_cost(rng, nx, 0)]
    Sm, sv = [None] * (N + 1), [None] * (N + 1)
    Kg, kg = [None] * N, [None] * N
    Sm[N], sv[N] = cost[N]["Q"], cost[N]["q"]
    for k in range(N - 1, -1, -1):
        A, B, b = dyn[k]["A"], dyn[k]["B"], dyn[k]["b"]
        Q, R, P, q, r = cost[k]["Q"], cost[k]["R"], cost[k]["S"], cost[k]["q"], cost[k]["r"]
        P_BTSmA = P + B.T @ Sm[k + 1] @ A
        invR = np.linalg.inv(R + B.T @ Sm[k + 1] @ B)
        rr = r + B.T @ sv[k + 1] + B.T @ Sm[k + 1] @ b
        Sm[k] = Q + A.T @ Sm[k + 1] @ A - P_BTSmA.T @ invR @ P_BTSmA
        sv[k] = q + A.T @ sv[k + 1] + A.T @ Sm[k + 1] @ b - P_BTSmA.T @ invR @ rr
        Kg[k] = -invR @ P_BTSmA
        kg[k] = -invR @ rr
    sol = orc.riccati(*stack(dyn, cost, nu)[:8], x0)
    assert np.allclose(sol["P"], np.stack(Sm), atol=TOL, rtol=0)
    assert np.allclose(sol["p"], np.stack(sv), atol=TOL, rtol=0)
    assert np.allclose(sol["K"], np.stack(Kg), atol=TOL, rtol=0)
    assert np.allclose(sol["k"], np.stack(kg), atol=TOL, rtol=0)
    for k in range(N):
        assert np.allclose(sol["du"][k], sol["K"][k] @ sol["dx"][k] + sol["k"][k], atol=1e-12)
    assert np.allclose(sol["dx"][0], x0)


def _known_solution_problem(rng, nus, nx=3):
    N = len(nus)
    xs = [rng.uniform(-1, 1, nx)]
    us, dyn, cost = [], [], []
    for k in range(N):
        m = nus[k]
        us.append(rng.uniform(-1, 1, m))
        d = rand_dyn(rng, nx, m)
        dyn.append(d)
        xs.append(d["b"] + d["A"] @ xs[k] + d["B"] @ us[k])
        c = rand_cost(rng, nx, m)
        c["q"] = -(c["Q"] @ xs[k] + c["S"].T @ us[k])
        c["r"] = -(c["R"] @ us[k] + c["S"] @ xs[k])
        cost.append(c)
    c = rand_cost(rng, nx, 0)
    c["q"] = -c["Q"] @ xs[N]
    cost.append(c)
    return xs, us, dyn, cost


def test_known_solution():
    """testHpipmInterface.cpp:112-152 knownSolution"""
    rng = np.random.default_rng(10)
    xs, us, dyn, cost = _known_solution_problem(rng, [2] * 5)
    A, B, b, Q, S, R, q, r, nu = stack(dyn, cost, 2)
    sol = orc.riccati(A, B, b, Q, S, R, q, r, xs[0], nu)
    assert np.allclose(sol["dx"], np.stack(xs), atol=TOL, rtol=0)
    assert np.allclose(sol["du"], np.stack(us), atol=TOL, rtol=0)


def test_no_inputs_stage():
    """testHpipmInterface.cpp:208-256 noInputs: stage 1 has nu = 0 (the event-node shape)"""
    rng = np.random.default_rng(11)
    nus = [2, 0, 2, 2, 2]
    xs, us, dyn, cost = _known_solution_problem(rng, nus)
    A, B, b, Q, S, R, q, r, nu = stack(dyn, cost, 2)
    sol = orc.riccati(A, B, b, Q, S, R, q, r, xs[0], nu)
    assert np.allclose(sol["dx"], np.stack(xs), atol=TOL, rtol=0)
    for k in range(5):
        assert np.allclose(sol["du"][k, : nus[k]], us[k], atol=TOL, rtol=0)


def test_dynamic_feasibility():
    """testHpipmInterface.cpp:75-110: x_{k+1} = A x + B u + b along the QP solution"""
    rng = np.random.default_rng(12)
    nx, nu, N = 3, 2, 5
    dyn = [rand_dyn(rng, nx, nu) for _ in range(N)]
    cost = [rand_cost(rng, nx, nu) for _ in range(N)] + [rand_cost(rng, nx, 0)]
    x0 = rng.uniform(-1, 1, nx)
    sol = orc.riccati(*stack(dyn, cost, nu)[:8], x0)
    for k in range(N):
        assert np.allclose(sol["dx"][k + 1], dyn[k]["A"] @ sol["dx"][k] + dyn[k]["B"] @ sol["du"][k] + dyn[k]["b"], atol=1e-12)


def test_kkt_residual_dense():
    """OcpToKkt-style check (ocs2_oc/test/oc_problem/testOcpToKkt.cpp:69-94): the Riccati solution zeroes the dense KKT system."""
    rng = np.random.default_rng(13)
    nx, nu, N = 4, 3, 6
    dyn = [rand_dyn(rng, nx, nu) for _ in range(N)]
    cost = [rand_cost(rng, nx, nu) for _ in range(N)] + [rand_cost(rng, nx, 0)]
    x0 = rng.uniform(-1, 1, nx)
    sol = orc.riccati(*stack(dyn, cost, nu)[:8], x0, reg=0.0)
    # costates from the value function: lam_k = P_k dx_k + p_k
    lam = [sol["P"][k] @ sol["dx"][k] + sol["p"][k] for k in range(N + 1)]
    for k in range(N):
        c, d = cost[k], dyn[k]
        gu = c["R"] @ sol["du"][k] + c["S"] @ sol["dx"][k] + c["r"] + d["B"].T @ lam[k + 1]
        assert np.max(np.abs(gu)) < 1e-9
        if k > 0:
            gx = c["Q"] @ sol["dx"][k] + c["S"].T @ sol["du"][k] + c["q"] + d["A"].T @ lam[k + 1] - lam[k]
            assert np.max(np.abs(gx)) < 1e-9
    gx = cost[N]["Q"] @ sol["dx"][N] + cost[N]["q"] - lam[N]
    assert np.max(np.abs(gx)) < 1e-9


def test_lu_projection_properties():
    """ocs2_core/test/misc/testLinearAlgebra.cpp:69-109: D Pu = 0, D Px = -C, D u0 = -e, Pu full column rank"""
    rng = np.random.default_rng(14)
    for nc, nx, nu in [(2, 3, 4), (12, 58, 35), (14, 58, 35), (1, 2, 2)]:
        Cm, D, e = rng.uniform(-1, 1, (nc, nx)), rng.uniform(-1, 1, (nc, nu)), rng.uniform(-1, 1, nc)
        Pu, Px, u0, rank = orc.lu_projection(Cm, D, e)
        assert rank == nc
        assert np.allclose(D @ Pu, 0, atol=1e-10)
        assert np.allclose(D @ Px, -Cm, atol=1e-10)
        assert np.allclose(D @ u0, -e, atol=1e-10)
        assert np.linalg.matrix_rank(Pu) == nu - nc


def test_change_of_input_variables():
    """ocs2_oc/test/testChangeOfInputVariables.cpp:68-257: projected model == original model evaluated at u = Pu ut + Px x + u0"""
    rng = np.random.default_rng(15)
    nx, nu, nc = 5, 4, 2
    d, c = rand_dyn(rng, nx, nu), rand_cost(rng, nx, nu)
    c0 = 0.3
    Cm, D, e = rng.uniform(-1, 1, (nc, nx)), rng.uniform(-1, 1, (nc, nu)), rng.uniform(-1, 1, nc)
    Pu, Px, u0, _ = orc.lu_projection(Cm, D, e)
    out = orc.change_of_input_variables(d["A"], d["B"], d["b"], c["Q"], c["S"], c["R"], c["q"], c["r"], c0, Pu, Px, u0)
    x, ut = rng.uniform(-1, 1, nx), rng.uniform(-1, 1, nu - nc)
    u = Pu @ ut + Px @ x + u0
    f = lambda x, u: c0 + c["q"] @ x + c["r"] @ u + 0.5 * x @ c["Q"] @ x + 0.5 * u @ c["R"] @ u + u @ c["S"] @ x
    ft = out["c"] + out["q"] @ x + out["r"] @ ut + 0.5 * x @ out["Q"] @ x + 0.5 * ut @ out["R"] @ ut + ut @ out["S"] @ x
    assert abs(f(x, u) - ft) < 1e-12
    assert np.allclose(d["A"] @ x + d["B"] @ u + d["b"], out["A"] @ x + out["B"] @ ut + out["b"], atol=1e-12)


def test_rk4_sensitivity_linear_system():
    """ocs2_core/test/integration/testSensitivityIntegrator.cpp:134-211: RK4 of x' = Ax + Bu has the closed-form
    A_d = sum_{j<=4} (A dt)^j / j!,  B_d = sum_{j<=3} (A dt)^j/(j+1)! dt B ; value-only path equals sensitivity path."""
    rng = np.random.default_rng(16)
    nx, nu, dt = 4, 2, 0.05
    A, B = rng.uniform(-1, 1, (nx, nx)), rng.uniform(-1, 1, (nx, nu))
    x, u = rng.uniform(-1, 1, nx), rng.uniform(-1, 1, nu)
    Ad, Bd, xn, xv = orc.rk4_sensitivity_linear(A, B, x, u, dt)
    Adt = A * dt
    I = np.eye(nx)
    Ad_ref = I + Adt + Adt @ Adt / 2 + Adt @ Adt @ Adt / 6 + Adt @ Adt @ Adt @ Adt / 24
    Bd_ref = (I + Adt / 2 + Adt @ Adt / 6 + Adt @ Adt @ Adt / 24) @ B * dt
    assert np.allclose(Ad, Ad_ref, atol=1e-14)
    assert np.allclose(Bd, Bd_ref, atol=1e-14)
    assert np.allclose(xn, Ad_ref @ x + Bd_ref @ u, atol=1e-14)
    assert np.allclose(xn, xv, atol=1e-15)


def test_time_discretization():
    """ocs2_oc/test/oc_data/testTimeDiscretization.cpp:36-129"""
    eps = 1e-9
    t, e = orc.time_discretization(0.1, 0.4, 0.1, [])
    assert np.allclose(t, [0.1, 0.2, 0.3, 0.4], atol=1e-15) and not e.any()
    # event inside: duplicated pre/post nodes
    t, e = orc.time_discretization(0.0, 0.3, 0.1, [0.15])
    assert np.allclose(t, [0.0, 0.1, 0.15, 0.15, 0.25, 0.3]) and list(e) == [0, 0, 1, 2, 0, 0]
    # event at the start is a PostEvent start; event at the end ignored
    t, e = orc.time_discretization(0.1, 0.3, 0.1, [0.1, 0.3])
    assert list(e) == [2, 0, 0] and np.allclose(t, [0.1, 0.2, 0.3])
    # an event too close to a grid point replaces it (dt_min merge)
    t, e = orc.time_discretization(0.0, 0.3, 0.1, [0.2 + 1e-7])
    assert len(t) == 5 and list(e) == [0, 0, 1, 2, 0]
    # horizon not a multiple of dt
    t, e = orc.time_discretization(0.0, 0.25, 0.1, [])
    assert np.allclose(t, [0.0, 0.1, 0.2, 0.25])
    del eps


def _lq_problem(rng, n=3, m=2):
    d, c = rand_dyn(rng, n, m), rand_cost(rng, n, m)
    return d, c


def test_sqp_unconstrained_lq_converges_in_one_step():
    """ocs2_sqp/test/testUnconstrained.cpp:97-162: on an LQ problem SQP needs <= 2 iterations, dynamics SSE < 1e-9,
    controller consistent with the primal solution."""
    rng = np.random.default_rng(20)
    n, m = 3, 2
    d, c = _lq_problem(rng)
    res = orc.sqp_test_problem(0, n, m, np.ones(n), 0.0, 1.0, 0.05, 10, A=d["A"], B=d["B"], Q=c["Q"], R=c["R"], P=c["S"], Qf=c["Q"],
                               xRef=np.ones(n), uRef=np.ones(m))
    assert len(res["log"]) <= 2
    assert res["log"][-1][6] < 1e-9  # dynamicsViolationSSE after the step
    assert res["log"][0][8] == 1.0   # full step accepted
    assert len(res["t"]) == 21
    # second iteration (if any) makes no progress: already optimal
    if len(res["log"]) == 2:
        assert res["log"][1][10] < 1e-6 and res["log"][1][11] < 1e-6


def test_sqp_switched_constraint_with_event():
    """ocs2_sqp/test/testSwitchedProblem.cpp:154-195: u[0]=0 before the event at 0.1875, u[1]=0 after; nodes 4,5 at the event."""
    rng = np.random.default_rng(21)
    n, m = 3, 2
    d = rand_dyn(rng, n, m)
    G = rng.uniform(-1, 1, (n, n))
    c1, c2, ce1, ce2, cf1, cf2 = (rand_cost(rng, n, m), rand_cost(rng, n, m), rand_cost(rng, n, 0), rand_cost(rng, n, 0),
                                  rand_cost(rng, n, 0), rand_cost(rng, n, 0))
    ev = 0.1875
    res = orc.sqp_test_problem(0, n, m, rng.uniform(-1, 1, n), 0.0, 1.0, 0.05, 20, A=d["A"], B=d["B"], G=G, Q=c1["Q"] + c2["Q"],
                               R=c1["R"] + c2["R"], P=c1["S"] + c2["S"], Qf=cf1["Q"] + cf2["Q"], Qe=ce1["Q"] + ce2["Q"],
                               xRef=rng.uniform(-1, 1, n), uRef=rng.uniform(-1, 1, m), Cm=np.zeros((2, n)), Dm=np.eye(2), em=np.zeros(2),
                               event_times=[ev], mode_sequence=[0, 1])
    assert len(res["log"]) <= 2
    assert res["log"][-1][6] < TOL
    assert res["t"][4] == ev and res["t"][5] == ev
    assert list(res["event"][4:6]) == [1, 2]
    assert abs(res["u"][3][0]) < TOL          # last node before the event: mode 0 -> u[0] = 0
    assert abs(res["u"][5][1]) < TOL          # first node after the event: mode 1 -> u[1] = 0
    for i in range(len(res["u"])):
        if res["event"][i] == 1:
            continue
        col = 0 if res["t"][i] < ev or (res["t"][i] == ev and res["event"][i] != 2) else 1
        assert abs(res["u"][i][col]) < TOL
        # the remapped feedback gain keeps the constraint: row `col` of K is zero (D K + C = 0)
        assert np.max(np.abs(res["K"][i][col])) < TOL


def test_sqp_event_at_beginning_and_end():
    """testSwitchedProblem.cpp:197-269"""
    rng = np.random.default_rng(22)
    n, m = 3, 2
    d, c = _lq_problem(rng)
    kw = dict(A=d["A"], B=d["B"], Q=c["Q"], R=c["R"], P=c["S"], Qf=c["Q"], xRef=np.zeros(n), uRef=np.zeros(m), Cm=np.zeros((2, n)),
              Dm=np.eye(2), em=np.zeros(2), mode_sequence=[0, 1])
    r0 = orc.sqp_test_problem(0, n, m, np.ones(n), 0.0, 1.0, 0.05, 20, event_times=[1e-8], **kw)
    assert r0["t"][0] == 1e-8 and r0["t"][1] != 1e-8 and r0["event"][0] == 2
    assert np.max(np.abs(r0["u"][:, 1])) < TOL
    r1 = orc.sqp_test_problem(0, n, m, np.ones(n), 0.0, 1.0, 0.05, 20, event_times=[1.0 - 1e-8], **kw)
    assert not np.any(r1["t"] == 1.0 - 1e-8)
    assert np.max(np.abs(r1["u"][:, 0])) < TOL


def test_sqp_circular_kinematics():
    """ocs2_sqp/test/testCircularKinematics.cpp:38-89 (projection on): nonlinear cost + projected constraint x.u = 0,
    SSE < 1e-6 and u == K-consistent feedforward."""
    res = orc.sqp_test_problem(1, 2, 2, np.array([1.0, 0.0]), 0.0, 1.0, 0.01, 20)
    last = res["log"][-1]
    assert last[6] < 1e-6 and last[7] < 1e-6
    assert np.allclose(res["x"][0], [1.0, 0.0])
    assert res["t"][0] == 0.0 and res["t"][-1] == 1.0
    # the particle stays on the unit circle and moves counter-clockwise at ~1 m/s
    rad = np.linalg.norm(res["x"], axis=1)
    assert np.max(np.abs(rad - 1.0)) < 1e-2
    assert res["x"][-1][1] > 0.5


def test_transcription_equals_metrics():
    """ocs2_oc/test/multiple_shooting/testTranscriptionPerformanceIndex.cpp:40-121: the performance index accumulated by the LQ
    transcription equals the value-only path at 1e-12 -- visible in the log as: baseline(iter k+1) == performanceAfterStep(iter k)."""
    res = orc.sqp_test_problem(1, 2, 2, np.array([1.0, 0.0]), 0.0, 1.0, 0.01, 6)
    log = res["log"]
    assert len(log) >= 3
    for k in range(len(log) - 1):
        assert np.allclose(log[k + 1][0:4], log[k][4:8], rtol=0, atol=1e-12)
