"""Centroidal flow map through the C ABI (b200sqp_centroidal_flow_map): the batched counterpart of PinocchioCentroidalDynamicsAD."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi
from . import lib as _l


def flow_map(model: dict, x, u, derivatives: bool = True, device: int = 0):
    """x [B, 12+nj], u [B, 12+nj] -> xdot [B, nx] (and dfdx [B, nx, nx], dfdu [B, nx, nu])"""
    x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
    u = np.ascontiguousarray(np.atleast_2d(u), dtype=np.float64)
    B, nx = x.shape
    desc = abi.model_desc(model)
    xd = np.zeros((B, nx))
    A = np.zeros((B, nx, nx)) if derivatives else None
    Bm = np.zeros((B, nx, nx)) if derivatives else None
    p = lambda a: None if a is None else a.ctypes.data_as(_l.dp)
    _l.check(_l.lib().b200sqp_centroidal_flow_map(C.byref(desc), C.c_int(B), p(x), p(u), p(xd), p(A), p(Bm), C.c_int(device)))
    if not derivatives:
        return xd
    return xd, np.swapaxes(A, 1, 2).copy(), np.swapaxes(Bm, 1, 2).copy()   # column-major -> [row, col]


def base_velocity(model: dict, x0, device: int = 0):
    """Ab^-1 * x0[:6]: the base velocity CentroidalMpcTargetTrajectoriesCalculator::commandedVelocityToTargetTrajectories derives from the
    NORMALIZED momentum (CentroidalMpcTargetTrajectoriesCalculator.cpp:121-125) = flow_map(x0, u = 0)[6:12] / mass.  x0 [nx] or [B, nx]."""
    x = np.atleast_2d(np.asarray(x0, float))
    if int(model.get("centroidalModelType", 0)) == 1:
        bv = np.array([srbd_base_velocity(model, xi) for xi in x])
        return bv[0] if np.ndim(x0) == 1 else bv
    xd = flow_map(model, x, np.zeros_like(x), derivatives=False, device=device)
    bv = xd[:, 6:12] / sum(model["mass"])
    return bv[0] if np.ndim(x0) == 1 else bv


def srbd_base_velocity(model: dict, x0):
    """Ab^-1 * x0[:6] for the SingleRigidBodyDynamics model, closed form (updateCentroidalDynamics, ModelHelperFunctions.cpp:61-79, and the block
    inverse of ModelHelperFunctionsImpl.h:40-47): Ab = [[m 1, m [R r]x T], [0, R I R' T]] with the nominal inertia I and com offset r."""
    z, y, xr = x0[9:12]
    cz, sz, cy, sy, cx, sx = np.cos(z), np.sin(z), np.cos(y), np.sin(y), np.cos(xr), np.sin(xr)
    R = np.array([[cz * cy, cz * sy * sx - sz * cx, cz * sy * cx + sz * sx], [sz * cy, sz * sy * sx + cz * cx, sz * sy * cx - cz * sx], [-sy, cy * sx, cy * cx]])
    T = R @ np.array([[-sy, 0, 1], [cy * sx, cx, 0], [cy * cx, -sx, 0.0]])   # Euler-rate (z, y, x) -> global angular velocity
    m = sum(model["mass"])
    rw = R @ np.asarray(model["srbd_nominal"]["com_to_base"])
    S = np.array([[0, -rw[2], rw[1]], [rw[2], 0, -rw[0]], [-rw[1], rw[0], 0]])
    A22 = R @ np.asarray(model["srbd_nominal"]["inertia"]) @ R.T @ T
    w = np.linalg.solve(A22, x0[3:6])
    v = (x0[0:3] - m * S @ T @ w) / m
    return np.concatenate([v, w])
