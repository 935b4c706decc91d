"""ctypes view of the C++ host layer (wb_humanoid_mpc_b200/host/*.hpp -> libb200sqp_host.so) for tests and bench.

The host layer itself is C++ (the reference's host code is C++): `b200sqp::host::SqpSolver` mirrors `ocs2::SqpSolver` over the C ABI and
`buildInstance` restates what `SolverBase::preRun` produces.  This module only builds the shared library (g++, links libb200sqp.so) and
exposes its `extern "C"` test entry points."""
from __future__ import annotations

import ctypes as C
import shutil
import subprocess
from pathlib import Path

import numpy as np

from . import abi
from . import lib as _l

PKG = Path(__file__).resolve().parent
HOST = PKG / "host"
SO = PKG / "libb200sqp_host.so"
MODEL_TXT = PKG / "data" / "g1_wb_model.txt"
CEN_MODEL_TXT = PKG / "data" / "g1_centroidal_model.txt"
_lib = None
dp = C.POINTER(C.c_double)
u8p = C.POINTER(C.c_uint8)


def sources():
    return sorted(HOST.glob("*.hpp")) + sorted(HOST.glob("*.cpp")) + [PKG.parent / "include" / "b200sqp.h"]


def build(force: bool = False) -> Path:
    core = _l.build()  # libb200sqp.so must exist to link against
    if not force and SO.exists() and all(s.stat().st_mtime <= SO.stat().st_mtime for s in sources() + [core]):
        return SO
    gxx = shutil.which("g++") or "g++"
    cmd = [gxx, "-std=c++17", "-O2", "-fPIC", "-shared", "-pthread", "-Wall", "-o", str(SO), str(HOST / "host_capi.cpp"), f"-L{PKG}", "-lb200sqp",
           "-Wl,-rpath,$ORIGIN"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stdout + res.stderr)
    return SO


def lib():
    global _lib
    if _lib is None:
        _l.lib()  # load libb200sqp.so first
        L = C.CDLL(str(build()))
        L.b200host_last_error.restype = C.c_char_p
        L.b200host_model_load.restype = C.c_void_p
        L.b200host_model_load.argtypes = [C.c_char_p]
        L.b200host_model_free.argtypes = [C.c_void_p]
        L.b200host_solver_create.restype = C.c_void_p
        L.b200host_solver_create.argtypes = [C.c_void_p, C.POINTER(abi.Settings), C.c_int, C.c_int, C.c_int]
        L.b200host_solver_destroy.argtypes = [C.c_void_p]
        L.b200host_build_instance.argtypes = [C.c_void_p, C.c_double, dp, C.c_double, C.c_char_p, C.c_double, dp, C.c_int, dp, dp, dp, C.c_int, dp, u8p, u8p,
                                              dp, dp, dp, dp, dp, dp, dp]
        L.b200host_solver_set_gait.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_double, C.c_double]
        L.b200host_solver_set_command.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, dp, dp, C.c_double, dp]
        L.b200host_centroidal_base_velocity.argtypes = [C.c_void_p, dp, dp]
        L.b200host_solver_reset.argtypes = [C.c_void_p]
        L.b200host_solver_run.argtypes = [C.c_void_p, C.c_void_p, C.c_double, dp, C.c_double]
        L.b200host_solver_n_nodes.argtypes = [C.c_void_p, C.c_int]
        L.b200host_solver_get_primal.argtypes = [C.c_void_p, C.c_int, dp, dp, dp]
        L.b200host_solver_get_log.argtypes = [C.c_void_p, C.c_int, dp, C.c_int]
        L.b200host_solver_benchmarks.argtypes = [C.c_void_p, dp]
        L.b200host_solver_value_function.argtypes = [C.c_void_p, C.c_int, C.c_double, dp, C.c_int, dp, dp]
        L.b200host_model_desc.argtypes = [C.c_void_p, C.POINTER(abi.ModelDesc), C.POINTER(abi.Settings)]
        L.b200host_model_dims.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), dp, dp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(dp)


def _check(rc):
    if rc is None or (isinstance(rc, int) and rc < 0):
        raise RuntimeError("b200sqp host: " + lib().b200host_last_error().decode())
    return rc


class HostModel:
    def __init__(self, path=MODEL_TXT, config=None, centroidal=False):
        """path: the flat model file; config = (urdf, task.info, reference.info, gait.info): the C++ loader of the reference's own files instead
        (centroidal: the centroidal MPC's task / reference files)"""
        L = lib()
        if config is not None:
            L.b200host_model_from_config.restype = C.c_void_p
            L.b200host_model_from_config.argtypes = [C.c_char_p] * 4 + [C.c_int]
            self.h = L.b200host_model_from_config(*[str(c).encode() for c in config], int(centroidal))
        else:
            self.h = L.b200host_model_load(str(path).encode())
        if not self.h:
            raise RuntimeError("b200sqp host: " + lib().b200host_last_error().decode())
        nx, nu, dt, hz = C.c_int(), C.c_int(), C.c_double(), C.c_double()
        lib().b200host_model_dims(self.h, C.byref(nx), C.byref(nu), C.byref(dt), C.byref(hz))
        self.nx, self.nu, self.dt, self.horizon = nx.value, nu.value, dt.value, hz.value

    def desc_and_settings(self):
        d, s = abi.ModelDesc(), abi.Settings()
        lib().b200host_model_desc(self.h, C.byref(d), C.byref(s))
        return d, s

    def dump(self):
        """everything else a HostModel holds, as one array (see b200host_model_dump)"""
        L = lib()
        L.b200host_model_dump.argtypes = [C.c_void_p, dp, C.c_int]
        n = L.b200host_model_dump(self.h, None, 0)
        out = np.zeros(n)
        L.b200host_model_dump(self.h, out.ctypes.data_as(dp), n)
        return out

    def cen_desc(self):
        """b200sqp_cen_desc of a centroidal model file, None for a whole-body one"""
        c = abi.CenDesc()
        L = lib()
        L.b200host_model_cen_desc.argtypes = [C.c_void_p, C.POINTER(abi.CenDesc)]
        return c if L.b200host_model_cen_desc(self.h, C.byref(c)) == 1 else None

    def base_velocity(self, x0):
        """Ab^-1 * x0[:6] of a centroidal state (device call unless the momentum is zero)"""
        x0, out = np.ascontiguousarray(x0, dtype=np.float64), np.zeros(6)
        _check(lib().b200host_centroidal_base_velocity(self.h, _p(x0), _p(out)))
        return out

    def build_instance(self, x0, t0=0.0, horizon=None, gait="stance", gait_start=None, cmd=None, previous=None, max_nodes=512, base_vel=None):
        horizon = self.horizon if horizon is None else horizon
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        n, nx, nu = max_nodes, self.nx, self.nu
        t, ev, cf = np.zeros(n), np.zeros(n, np.uint8), np.zeros((n, 2), np.uint8)
        sw, imp, arm, xr = np.zeros((n, 2, 3)), np.zeros((n, 2)), np.zeros(n), np.zeros((n, nx))
        xi, ui = np.zeros((n, nx)), np.zeros((n, nu))
        cm = None if cmd is None else np.ascontiguousarray(cmd, dtype=np.float64)
        if previous is not None:
            pt, px, pu = (np.ascontiguousarray(previous[k], dtype=np.float64) for k in ("t", "x", "u"))
            pn = len(pt)
        else:
            pt = px = pu = np.zeros(1)
            pn = 0
        k = _check(lib().b200host_build_instance(self.h, t0, _p(x0), horizon, gait.encode(), t0 if gait_start is None else gait_start,
                                                 None if cm is None else _p(cm), pn, _p(pt), _p(px), _p(pu), n, _p(t), ev.ctypes.data_as(u8p),
                                                 cf.ctypes.data_as(u8p), _p(sw), _p(imp), _p(arm), _p(xr), _p(xi), _p(ui),
                                                 None if base_vel is None else _p(np.ascontiguousarray(base_vel, dtype=np.float64))))
        return dict(x0=x0, x_init=xi[:k], u_init=ui[:k - 1], t_nodes=t[:k], node_event=ev[:k], contact_flags=cf[:k], swing_ref=sw[:k],
                    impact_factor=imp[:k], arm_phase=arm[:k], x_ref=xr[:k])

    def close(self):
        if self.h:
            lib().b200host_model_free(self.h)
            self.h = None


def trajectory_spread(old_events, old_modes, new_events, new_modes, t, x, u):
    """C++ trajectorySpread on flat arrays -> (t, x, u, will_truncate, will_spread)"""
    oe, ne = np.ascontiguousarray(old_events, dtype=np.float64), np.ascontiguousarray(new_events, dtype=np.float64)
    om, nm = np.ascontiguousarray(old_modes, dtype=np.int32), np.ascontiguousarray(new_modes, dtype=np.int32)
    t, x, u = (np.array(a, dtype=np.float64, order="C") for a in (t, x, u))
    fl = np.zeros(2, dtype=np.int32)
    ip = C.POINTER(C.c_int)
    L = lib()
    L.b200host_trajectory_spread.argtypes = [C.c_int, dp, ip, C.c_int, dp, ip, C.c_int, C.c_int, C.c_int, dp, dp, dp, ip]
    m = _check(L.b200host_trajectory_spread(len(oe), _p(oe), om.ctypes.data_as(ip), len(ne), _p(ne), nm.ctypes.data_as(ip), len(t), x.shape[1], u.shape[1],
                                            _p(t), _p(x), _p(u), fl.ctypes.data_as(ip)))
    return t[:m], x[:m], u[:m], bool(fl[0]), bool(fl[1])


class HostSqpSolver:
    """b200sqp::host::SqpSolver (C++) driven from Python: the reference-facing call path of bench.py's e2e number."""

    def __init__(self, model: HostModel, settings: abi.Settings, batch: int, device: int = 0, threads: int = 0):
        self.model, self.batch, self.settings = model, batch, settings
        self.h = lib().b200host_solver_create(model.h, C.byref(settings), batch, device, threads)
        if not self.h:
            raise RuntimeError("b200sqp host: " + lib().b200host_last_error().decode())

    def set_gait(self, b, gait, start, final):
        _check(lib().b200host_solver_set_gait(self.h, b, gait.encode(), start, final))

    def set_command(self, b, t0, x0, cmd, horizon, base_vel=None):
        x0, cmd = np.ascontiguousarray(x0, dtype=np.float64), np.ascontiguousarray(cmd, dtype=np.float64)
        bv = None if base_vel is None else _p(np.ascontiguousarray(base_vel, dtype=np.float64))
        _check(lib().b200host_solver_set_command(self.h, self.model.h, b, t0, _p(x0), _p(cmd), horizon, bv))

    def set_trajectory_spread(self, on: bool):
        L = lib()
        L.b200host_solver_set_trajectory_spread.argtypes = [C.c_void_p, C.c_int]
        L.b200host_solver_set_trajectory_spread(self.h, int(on))

    def set_exclusive_solve(self, on: bool):
        """double buffering with several solver objects: one solve on the GPU at a time (SqpSolver::setExclusiveSolve)"""
        L = lib()
        L.b200host_solver_set_exclusive_solve.argtypes = [C.c_void_p, C.c_int]
        L.b200host_solver_set_exclusive_solve(self.h, int(on))

    def reset(self):
        _check(lib().b200host_solver_reset(self.h))

    def run(self, t0, x0s, tf):
        x0s = np.ascontiguousarray(x0s, dtype=np.float64)
        assert x0s.shape == (self.batch, self.model.nx)
        _check(lib().b200host_solver_run(self.h, self.model.h, t0, _p(x0s), tf))

    def primal_solution(self, b):
        n = _check(lib().b200host_solver_n_nodes(self.h, b))
        t, x, u = np.zeros(n), np.zeros((n, self.model.nx)), np.zeros((n, self.model.nu))
        _check(lib().b200host_solver_get_primal(self.h, b, _p(t), _p(x), _p(u)))
        return dict(t=t, x=x, u=u)

    def iterations_log(self, b):
        out = np.zeros((self.settings.sqp_iteration, 12))
        k = _check(lib().b200host_solver_get_log(self.h, b, _p(out), self.settings.sqp_iteration))
        return out[:k]

    def write_log(self, time=0.0) -> str:
        """sqp::Logger CSV of the last run (host/SqpLogging.hpp)"""
        L = lib()
        L.b200host_solver_write_log.argtypes = [C.c_void_p, C.c_double, C.c_char_p, C.c_int]
        n = _check(L.b200host_solver_write_log(self.h, time, None, 0))
        buf = C.create_string_buffer(n + 1)
        _check(L.b200host_solver_write_log(self.h, time, buf, n + 1))
        return buf.value.decode()

    def benchmarks(self):
        """ms: [LQ, QP, line search, projection share | host preRun, pack, upload, solve, download, unpack]"""
        ms = np.zeros(10)
        lib().b200host_solver_benchmarks(self.h, _p(ms))
        return ms

    def value_function(self, b, t, x):
        nx = self.model.nx
        x = np.ascontiguousarray(x, dtype=np.float64)
        P, p = np.zeros((nx, nx)), np.zeros(nx)
        _check(lib().b200host_solver_value_function(self.h, b, t, _p(x), nx, _p(P), _p(p)))
        return P.T.copy(), p

    def close(self):
        if self.h:
            lib().b200host_solver_destroy(self.h)
            self.h = None
