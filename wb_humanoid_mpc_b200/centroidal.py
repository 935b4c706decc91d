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
    xd = flow_map(model, x, np.zeros_like(x), derivatives=False, device=device)
    bv = xd[:, 6:12] / sum(model["mass"])
    return bv[0] if np.ndim(x0) == 1 else bv
