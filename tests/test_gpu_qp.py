"""K2 parity (GPU, through the C ABI): batched Riccati vs the CPU oracle and vs the reference's known-answer recipes."""
import numpy as np
import pytest

import oracle_lib as orc
from test_oracle_qp import _known_solution_problem, rand_cost, rand_dyn, stack

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["cluster", "two_kernel", "generic"])
def k2_mode(request, monkeypatch):
    """the three K2 variants for the whole-body sizes (b200sqp.cu launch_riccati): 4-CTA cluster kernel (small batches), riccati_wb.cuh's backward
    + forward kernels (the large-batch default), the generic one-CTA kernel"""
    monkeypatch.setenv("B200SQP_NO_CLUSTER", "0" if request.param == "cluster" else "1")
    monkeypatch.setenv("B200SQP_K2_LEGACY", "1" if request.param == "generic" else "0")
    return request.param


def _random_batch(rng, Bn, N, nx, numax, nu_pattern=None):
    A = rng.uniform(-1, 1, (Bn, N, nx, nx)) * (0.6 / np.sqrt(nx))
    A += np.eye(nx)
    Bm = rng.uniform(-1, 1, (Bn, N, nx, numax)) * 0.3
    b = rng.uniform(-1, 1, (Bn, N, nx)) * 0.1
    M = rng.uniform(-1, 1, (Bn, N + 1, nx + numax, nx + numax))
    H = np.einsum("bkij,bkil->bkjl", M, M) / (nx + numax) + 0.05 * np.eye(nx + numax)
    Q, S, R = H[:, :, :nx, :nx], H[:, :N, nx:, :nx], H[:, :N, nx:, nx:]
    q = rng.uniform(-1, 1, (Bn, N + 1, nx))
    r = rng.uniform(-1, 1, (Bn, N, numax))
    nu = np.full((Bn, N), numax, dtype=np.int32)
    if nu_pattern is not None:
        nu[:] = nu_pattern
    dx0 = rng.uniform(-1, 1, (Bn, nx)) * 0.1
    return [np.ascontiguousarray(a) for a in (A, Bm, b, Q, S, R, q, r)], nu, dx0


def _rel(a, b):
    return np.max(np.abs(a - b)) / max(1.0, np.max(np.abs(b)))


def _compare(sol, Bn, data, nu, dx0, tol=1e-9):
    A, Bm, b, Q, S, R, q, r = data
    for i in range(Bn):
        ref = orc.riccati(A[i], Bm[i], b[i], Q[i], S[i], R[i], q[i], r[i], dx0[i], nu[i])
        for key in ("dx", "du", "K", "k", "P", "p"):
            assert _rel(sol[key][i], ref[key]) < tol, (i, key, _rel(sol[key][i], ref[key]))


def test_riccati_small_known_solution():
    """testHpipmInterface.cpp knownSolution + noInputs recipes on the GPU path (nx=3, nu=2, N=5, stage 1 without inputs)."""
    from wb_humanoid_mpc_b200.qp import BatchedQp

    rng = np.random.default_rng(5)
    nus = [2, 0, 2, 2, 2]
    xs, us, dyn, cost = _known_solution_problem(rng, nus)
    A, Bm, b, Q, S, R, q, r, nu = stack(dyn, cost, 2)
    qp = BatchedQp(1, 5, 3, 2)
    qp.upload(A[None], Bm[None], b[None], Q[None], S[None], R[None], q[None], r[None], xs[0][None], nu[None])
    qp.solve()
    sol = qp.download()
    assert np.allclose(sol["dx"][0], np.stack(xs), atol=1e-9, rtol=0)
    for k in range(5):
        assert np.allclose(sol["du"][0, k, : nus[k]], us[k], atol=1e-9, rtol=0)


@pytest.mark.parametrize("nx,numax,N,Bn", [(3, 2, 5, 4), (35, 23, 20, 3), (58, 23, 100, 4)])
def test_riccati_vs_oracle(nx, numax, N, Bn):
    from wb_humanoid_mpc_b200.qp import BatchedQp

    rng = np.random.default_rng(nx * 1000 + N)
    data, nu, dx0 = _random_batch(rng, Bn, N, nx, numax)
    qp = BatchedQp(Bn, N, nx, numax)
    qp.upload(*data, dx0, nu)
    qp.solve()
    sol = qp.download()
    assert not sol["status"].any()
    _compare(sol, Bn, data, nu, dx0)


def test_riccati_whole_body_sizes_all_variants(k2_mode):
    from wb_humanoid_mpc_b200.qp import BatchedQp

    rng = np.random.default_rng(58023)
    nx, numax, N, Bn = 58, 23, 100, 4
    data, nu, dx0 = _random_batch(rng, Bn, N, nx, numax)
    qp = BatchedQp(Bn, N, nx, numax)
    qp.upload(*data, dx0, nu)
    qp.solve()
    sol = qp.download()
    assert not sol["status"].any()
    _compare(sol, Bn, data, nu, dx0)


def test_riccati_large_batch_default_dispatch():
    """a batch beyond the cluster threshold takes the two-kernel path by default; two waves of CTAs (batch > number of SMs); every instance
    against the oracle would take minutes, so instances 0, 17, 149, 299 are compared and all are checked for dynamic feasibility"""
    from wb_humanoid_mpc_b200.qp import BatchedQp

    rng = np.random.default_rng(300)
    nx, numax, N, Bn = 58, 23, 12, 300
    pattern = rng.choice([21, 22, 23, 23, 0], size=(Bn, N)).astype(np.int32)
    data, nu, dx0 = _random_batch(rng, Bn, N, nx, numax, pattern)
    qp = BatchedQp(Bn, N, nx, numax)
    qp.upload(*data, dx0, nu)
    qp.solve()
    sol = qp.download()
    assert not sol["status"].any()
    A, Bm, b, Q, S, R, q, r = data
    for i in (0, 17, 149, 299):
        ref = orc.riccati(A[i], Bm[i], b[i], Q[i], S[i], R[i], q[i], r[i], dx0[i], nu[i])
        for key in ("dx", "du", "K", "k", "P", "p"):
            assert _rel(sol[key][i], ref[key]) < 1e-9, (i, key)
    for i in range(Bn):
        for k in range(N):
            m = nu[i, k]
            xn = A[i, k] @ sol["dx"][i, k] + Bm[i, k][:, :m] @ sol["du"][i, k, :m] + b[i, k]
            assert np.max(np.abs(xn - sol["dx"][i, k + 1])) < 1e-10, (i, k)


def test_riccati_varying_nu_and_event_stages(k2_mode):
    """per-stage projected input dimension 21/22/23 and nu=0 event stages (G1 whole-body shapes)"""
    from wb_humanoid_mpc_b200.qp import BatchedQp

    rng = np.random.default_rng(77)
    nx, numax, N, Bn = 58, 23, 40, 3
    pattern = rng.choice([21, 22, 23, 23, 23, 0], size=(Bn, N)).astype(np.int32)
    data, nu, dx0 = _random_batch(rng, Bn, N, nx, numax, pattern)
    # event-node shape: A = I, no cost
    A, Bm, b, Q, S, R, q, r = data
    for i in range(Bn):
        for k in range(N):
            if nu[i, k] == 0:
                A[i, k] = np.eye(nx)
                Q[i, k] = 0
                q[i, k] = 0
    qp = BatchedQp(Bn, N, nx, numax)
    qp.upload(*data, dx0, nu)
    qp.solve()
    sol = qp.download()
    _compare(sol, Bn, data, nu, dx0)
    # feasibility + KKT stationarity from the value function (size-independent property)
    for i in range(Bn):
        lam = [sol["P"][i, k] @ sol["dx"][i, k] + sol["p"][i, k] for k in range(N + 1)]
        for k in range(N):
            m = nu[i, k]
            xn = A[i, k] @ sol["dx"][i, k] + Bm[i, k][:, :m] @ sol["du"][i, k, :m] + b[i, k]
            assert np.max(np.abs(xn - sol["dx"][i, k + 1])) < 1e-10
            if m:
                gu = R[i, k][:m, :m] @ sol["du"][i, k, :m] + S[i, k][:m] @ sol["dx"][i, k] + r[i, k, :m] + Bm[i, k][:, :m].T @ lam[k + 1]
                assert np.max(np.abs(gu)) < 1e-8


def test_riccati_failure_is_reported():
    """indefinite R -> hpipm status != SUCCESS -> '[SqpSolver] Failed to solve QP' (SqpSolver.cpp:306-308)"""
    from wb_humanoid_mpc_b200.lib import B200SqpError
    from wb_humanoid_mpc_b200.qp import BatchedQp

    rng = np.random.default_rng(3)
    data, nu, dx0 = _random_batch(rng, 2, 4, 3, 2)
    data[5][1] = -np.eye(2) * 100.0  # R of instance 1
    qp = BatchedQp(2, 4, 3, 2)
    qp.upload(*data, dx0, nu)
    qp.solve()
    with pytest.raises(B200SqpError) as ei:
        qp.download()
    assert ei.value.code == -4


def test_kkt_residual_full_size(k2_mode):
    """north_star parity list: the KKT residual of the GPU QP solution, assembled stage-wise as in OcpToKkt (no oracle involved):
    with costates lambda_k = P_k dx_k + p_k,
        stationarity in dx_k : Q dx + S' du + q + A' lambda_{k+1} - lambda_k = 0,   in du_k : S dx + R du + r + B' lambda_{k+1} = 0,
        dynamics            : dx_{k+1} = A dx + B du + b,   dx_0 given."""
    from wb_humanoid_mpc_b200.qp import BatchedQp

    rng = np.random.default_rng(21)
    nx, numax, N, Bn = 58, 23, 115, 3
    pattern = np.array([23 if k % 7 else 21 for k in range(N)], dtype=np.int32)
    data, nu, dx0 = _random_batch(rng, Bn, N, nx, numax, nu_pattern=pattern)
    A, Bm, b, Q, S, R, q, r = data
    qp = BatchedQp(Bn, N, nx, numax)
    qp.upload(*data, dx0, nu)
    qp.solve()
    sol = qp.download()
    for i in range(Bn):
        dx, du, P, p = sol["dx"][i], sol["du"][i], sol["P"][i], sol["p"][i]
        lam = np.einsum("kij,kj->ki", P, dx) + p
        assert np.allclose(dx[0], dx0[i], atol=1e-13)
        worst = 0.0
        for k in range(N):
            m = nu[i, k]
            uk = du[k, :m]
            Bk, Sk, Rk = Bm[i, k][:, :m], S[i, k][:m, :], R[i, k][:m, :m]
            dyn = A[i, k] @ dx[k] + Bk @ uk + b[i, k] - dx[k + 1]
            sx = Q[i, k] @ dx[k] + Sk.T @ uk + q[i, k] + A[i, k].T @ lam[k + 1] - lam[k]
            su = Sk @ dx[k] + Rk @ uk + r[i, k][:m] + Bk.T @ lam[k + 1]
            scale = max(1.0, np.abs(lam[k]).max())
            worst = max(worst, np.abs(dyn).max(), np.abs(sx).max() / scale, np.abs(su).max() / scale)
        term = Q[i, N] @ dx[N] + q[i, N] - lam[N]
        worst = max(worst, np.abs(term).max() / max(1.0, np.abs(lam[N]).max()))
        assert worst < 1e-9, worst


def test_cluster_variant_is_bitwise_identical_and_deterministic():
    """K2 dispatches small batches to the cluster variant (4 SMs per instance, exchanges through distributed shared memory).  It performs the
    same tile arithmetic as the one-CTA kernel, so every output must agree BIT FOR BIT, on every repetition (a lost or reordered remote store
    would show up as a difference)."""
    import os

    from wb_humanoid_mpc_b200.qp import BatchedQp

    rng = np.random.default_rng(33)
    nx, numax, N, Bn = 58, 23, 115, 8
    pattern = np.array([23 if k % 5 else (0 if k % 10 == 0 else 21) for k in range(N)], dtype=np.int32)   # includes nu = 0 (event) stages
    data, nu, dx0 = _random_batch(rng, Bn, N, nx, numax, nu_pattern=pattern)
    qp = BatchedQp(Bn, N, nx, numax)
    qp.upload(*data, dx0, nu)
    old = os.environ.get("B200SQP_NO_CLUSTER")
    old_legacy = os.environ.get("B200SQP_K2_LEGACY")
    try:
        os.environ["B200SQP_NO_CLUSTER"] = "1"
        os.environ["B200SQP_K2_LEGACY"] = "1"   # the generic one-CTA kernel is the bit-exact reference of the cluster variant
        qp.solve()
        ref = qp.download()
        os.environ["B200SQP_NO_CLUSTER"] = "0"
        for rep in range(25):
            qp.solve()
            sol = qp.download()
            for key in ("dx", "du", "K", "k", "P", "p"):
                assert np.array_equal(sol[key], ref[key]), (rep, key, np.abs(sol[key] - ref[key]).max())
    finally:
        if old_legacy is None:
            os.environ.pop("B200SQP_K2_LEGACY", None)
        else:
            os.environ["B200SQP_K2_LEGACY"] = old_legacy
        if old is None:
            os.environ.pop("B200SQP_NO_CLUSTER", None)
        else:
            os.environ["B200SQP_NO_CLUSTER"] = old
