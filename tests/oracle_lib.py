"""ctypes bindings of the CPU oracle (oracle/liboracle.so). Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
ORACLE_DIR = ROOT / "oracle"
_lib = None

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


def _p(a):
    return None if a is None else a.ctypes.data_as(dp)


def _pi(a):
    return None if a is None else a.ctypes.data_as(ip)


def F(a):
    """contiguous float64 copy"""
    return np.ascontiguousarray(a, dtype=np.float64)


def build_oracle(force=False):
    so = ORACLE_DIR / "liboracle.so"
    srcs = list(ORACLE_DIR.glob("*.hpp")) + list(ORACLE_DIR.glob("*.cpp")) + list(ORACLE_DIR.glob("*.inc"))
    if force or not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-C", str(ORACLE_DIR), "-s", "liboracle.so"], check=True, stdout=sys.stderr)
    return so


_fast = None


def build_fast(force=False):
    """oracle/fast/liboracle_fast.so: the fast CPU baseline (timed arm of bench.py); see oracle/fast/wb_fast.cu"""
    so = ORACLE_DIR / "fast" / "liboracle_fast.so"
    csrc = ROOT / "wb_humanoid_mpc_b200" / "csrc"
    srcs = [ORACLE_DIR / "fast" / "wb_fast.cu"] + list(csrc.glob("*.cuh")) + list(csrc.glob("*.inc"))
    stale = not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs)
    import shutil
    if (force or stale) and (shutil.which("nvcc") or (Path("/usr/local/cuda/bin/nvcc").exists())):
        subprocess.run(["make", "-C", str(ORACLE_DIR), "-s", "fast"] + (["-B"] if force else []), check=True, stdout=sys.stderr)
    return so


def fast_lib():
    global _fast
    if _fast is None:
        _fast = C.CDLL(str(build_fast()))
        _fast.orc_fast_wb_sqp_batch.restype = C.c_int
    return _fast


def fast_wb_sqp_batch(model, batch, settings, threads=1, node_threads=1, want_gains=False):
    """the fast CPU baseline on a stacked batch (solver.stack_instances layout) -> dict(x, u, log, n_iter, K, stage_s, seconds)"""
    import time

    from wb_humanoid_mpc_b200 import abi

    desc = abi.model_desc(model)
    u8 = lambda a: np.ascontiguousarray(a, dtype=np.uint8)
    B, n = batch["t_nodes"].shape
    nx, nu = model["nx"], model["nu"]
    x, u = F(batch["x_init"]).copy(), F(batch["u_init"]).copy()
    arrs = [F(batch["t_nodes"]), u8(batch["node_event"]), F(batch["x0"]), u8(batch["contact_flags"]), F(batch["swing_ref"]), F(batch["impact_factor"]),
            F(batch["arm_phase"]), F(batch["x_ref"])]
    log = np.zeros((B, settings.sqp_iteration, 16))
    n_iter = np.zeros(B, dtype=np.int32)
    K = np.zeros((B, n - 1, nx, nu)) if want_gains else None
    stage = np.zeros(3)
    u8p = C.POINTER(C.c_uint8)
    t0 = time.perf_counter()
    rc = fast_lib().orc_fast_wb_sqp_batch(C.byref(desc), C.c_int(B), C.c_int(threads), C.c_int(node_threads), C.c_int(n), _p(arrs[0]),
                                          arrs[1].ctypes.data_as(u8p), _p(arrs[2]), _p(x), _p(u), arrs[3].ctypes.data_as(u8p), _p(arrs[4]), _p(arrs[5]),
                                          _p(arrs[6]), _p(arrs[7]), C.byref(settings), _p(log), n_iter.ctypes.data_as(C.POINTER(C.c_int32)), _p(K), _p(stage))
    dt = time.perf_counter() - t0
    if rc != 0:
        raise RuntimeError(f"fast CPU baseline failed ({rc})")
    return dict(x=x, u=u, log=log, n_iter=n_iter, K=None if K is None else np.swapaxes(K, -1, -2).copy(), stage_s=stage, seconds=dt)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build_oracle()))
        _lib.orc_time_discretization.restype = C.c_int
        _lib.orc_riccati.restype = C.c_int
        _lib.orc_lu_projection.restype = C.c_int
        _lib.orc_sqp_test_problem.restype = C.c_int
    return _lib


def time_discretization(t0, tf, dt, events):
    ev = F(events)
    cap = int((tf - t0) / dt) * 2 + 2 * len(ev) + 16
    t = np.zeros(cap)
    e = np.zeros(cap, dtype=np.int32)
    n = lib().orc_time_discretization(C.c_double(t0), C.c_double(tf), C.c_double(dt), _p(ev), C.c_int(len(ev)), _p(t), _pi(e), C.c_int(cap))
    assert n <= cap
    return t[:n].copy(), e[:n].copy()


def trajectory_spread(old_events, old_modes, new_events, new_modes, t, x, tags, event_data=()):
    """oracle/trajectory_spreading.hpp on flat arrays -> dict(t, x, tags, will_truncate, will_spread, post_event_indices, event_data)"""
    oe, ne = F(old_events), F(new_events)
    om, nm = np.ascontiguousarray(old_modes, dtype=np.int32), np.ascontiguousarray(new_modes, dtype=np.int32)
    t, x = np.array(t, dtype=np.float64, order="C"), np.array(x, dtype=np.float64, order="C")
    tags = np.array(tags, dtype=np.int32, order="C")
    fl, post, npost = np.zeros(2, dtype=np.int32), np.zeros(len(t) + 1, dtype=np.int32), C.c_int(0)
    ev_in = F(event_data)
    ev_out, nev = np.zeros(len(ev_in) + 1), C.c_int(0)
    L = lib()
    L.orc_trajectory_spread.restype = C.c_int
    m = L.orc_trajectory_spread(C.c_int(len(oe)), _p(oe), _pi(om), C.c_int(len(ne)), _p(ne), _pi(nm), C.c_int(len(t)), C.c_int(x.shape[1]), _p(t), _p(x),
                                _pi(tags), _pi(fl), _pi(post), C.byref(npost), C.c_int(len(ev_in)), _p(ev_in), _p(ev_out), C.byref(nev))
    return dict(t=t[:m], x=x[:m], tags=tags[:m], will_truncate=bool(fl[0]), will_spread=bool(fl[1]), post_event_indices=post[: npost.value].tolist(),
                event_data=ev_out[: nev.value])


def riccati(A, B, b, Q, S, R, q, r, dx0, nu=None, reg=1e-12):
    """Padded stage arrays: A (N,nx,nx) etc. given as numpy arrays in math layout [k, row, col]."""
    N, nx = A.shape[0], A.shape[1]
    numax = B.shape[2]
    nu = np.full(N, numax, dtype=np.int32) if nu is None else np.ascontiguousarray(nu, dtype=np.int32)
    cm = lambda M: F(np.swapaxes(M, -1, -2))  # column-major per stage
    dx = np.zeros((N + 1, nx))
    du = np.zeros((N, numax))
    P = np.zeros((N + 1, nx, nx))
    p = np.zeros((N + 1, nx))
    K = np.zeros((N, nx, numax))
    kff = np.zeros((N, numax))
    rc = lib().orc_riccati(C.c_int(N), C.c_int(nx), C.c_int(numax), _pi(nu), _p(cm(A)), _p(cm(B)), _p(F(b)), _p(cm(Q)), _p(cm(S)),
                           _p(cm(R)), _p(F(q)), _p(F(r)), _p(F(dx0)), C.c_double(reg), _p(dx), _p(du), _p(P), _p(p), _p(K), _p(kff))
    if rc != 0:
        raise RuntimeError("oracle riccati failed")
    return dict(dx=dx, du=du, P=np.swapaxes(P, 1, 2).copy(), p=p, K=np.swapaxes(K, 1, 2).copy(), k=kff)


def lu_projection(Cm, D, e):
    nc, nx = Cm.shape
    nu = D.shape[1]
    Pu = np.zeros((nu - nc, nu))
    Px = np.zeros((nx, nu))
    u0 = np.zeros(nu)
    rank = lib().orc_lu_projection(C.c_int(nc), C.c_int(nx), C.c_int(nu), _p(F(Cm.T)), _p(F(D.T)), _p(F(e)), _p(Pu), _p(Px), _p(u0))
    return Pu.T.copy(), Px.T.copy(), u0, rank


def change_of_input_variables(A, B, b, Q, S, R, q, r, c, Pu, Px, u0):
    nx, nu = B.shape
    nut = Pu.shape[1]
    cmA, cmB, cmQ, cmS, cmR = F(A.T), F(B.T), F(Q.T), F(S.T), F(R.T)
    bb, qq, rr = F(b).copy(), F(q).copy(), F(r).copy()
    cc = C.c_double(c)
    Bt, St, Rt, rt = np.zeros((nut, nx)), np.zeros((nx, nut)), np.zeros((nut, nut)), np.zeros(nut)
    lib().orc_change_of_input_variables(C.c_int(nx), C.c_int(nu), C.c_int(nut), _p(cmA), _p(cmB), _p(bb), _p(cmQ), _p(cmS), _p(cmR), _p(qq),
                                        _p(rr), C.byref(cc), _p(F(Pu.T)), _p(F(Px.T)), _p(F(u0)), _p(Bt), _p(St), _p(Rt), _p(rt))
    return dict(A=cmA.T.copy(), B=Bt.T.copy(), b=bb, Q=cmQ.T.copy(), S=St.T.copy(), R=Rt.T.copy(), q=qq, r=rt, c=cc.value)


def rk4_sensitivity_linear(A, B, x, u, dt):
    nx, nu = B.shape
    Ad, Bd, xn, xv = np.zeros((nx, nx)), np.zeros((nu, nx)), np.zeros(nx), np.zeros(nx)
    lib().orc_rk4_sensitivity_linear(C.c_int(nx), C.c_int(nu), _p(F(A.T)), _p(F(B.T)), _p(F(x)), _p(F(u)), C.c_double(dt), _p(Ad), _p(Bd),
                                     _p(xn), _p(xv))
    return Ad.T.copy(), Bd.T.copy(), xn, xv


class _LqProblem(C.Structure):
    _fields_ = [("kind", C.c_int), ("nx", C.c_int), ("nu", C.c_int), ("A", dp), ("B", dp), ("G", dp), ("Q", dp), ("R", dp), ("P", dp),
                ("Qf", dp), ("Qe", dp), ("xRef", dp), ("uRef", dp), ("nModes", C.c_int), ("Cm", dp), ("Dm", dp), ("em", dp),
                ("nEvents", C.c_int), ("eventTimes", dp), ("modeSequence", ip), ("t0", C.c_double), ("tf", C.c_double), ("x0", dp),
                ("dt", C.c_double), ("sqpIteration", C.c_int)]


def sqp_test_problem(kind, nx, nu, x0, t0, tf, dt, sqp_iteration, A=None, B=None, G=None, Q=None, R=None, P=None, Qf=None, Qe=None, xRef=None,
                     uRef=None, Cm=None, Dm=None, em=None, event_times=(), mode_sequence=None):
    keep = []

    def cm(M):
        if M is None:
            return None
        a = F(np.asarray(M).T)
        keep.append(a)
        return _p(a)

    def v(x):
        if x is None:
            return None
        a = F(x)
        keep.append(a)
        return _p(a)

    ev = F(event_times)
    ms = np.ascontiguousarray(mode_sequence if mode_sequence is not None else [-1] * (len(ev) + 1), dtype=np.int32)
    nmodes = 0 if Cm is None else np.asarray(Cm).shape[0]
    pr = _LqProblem(kind, nx, nu, cm(A), cm(B), cm(G), cm(Q), cm(R), cm(P), cm(Qf), cm(Qe), v(xRef), v(uRef), nmodes,
                    v(None if Cm is None else np.asarray(Cm)), v(None if Dm is None else np.asarray(Dm)), v(em), len(ev), _p(ev), _pi(ms),
                    t0, tf, v(x0), dt, sqp_iteration)
    cap = int((tf - t0) / dt) * 2 + 2 * len(ev) + 16
    times, events = np.zeros(cap), np.zeros(cap, dtype=np.int32)
    x, u, K = np.zeros((cap, nx)), np.zeros((cap, nu)), np.zeros((cap, nx, nu))
    log = np.zeros((64, 16))
    nit = C.c_int(0)
    n = lib().orc_sqp_test_problem(C.byref(pr), C.c_int(cap), _p(times), _pi(events), _p(x), _p(u), _p(K), C.c_int(64), _p(log), C.byref(nit))
    assert n > 0, n
    return dict(t=times[:n], event=events[:n], x=x[:n], u=u[: n - 1], K=np.swapaxes(K[: n - 1], 1, 2).copy(), log=log[: nit.value])


LOG_FIELDS = ["base_merit", "base_cost", "base_dynSSE", "base_eqSSE", "merit", "cost", "dynSSE", "eqSSE", "stepSize", "stepType", "dx_norm",
              "du_norm", "armijo", "convergence"]


# ------------------------------------------------------------------------------------------------------------------
# whole-body restatement
# ------------------------------------------------------------------------------------------------------------------
u8p = C.POINTER(C.c_uint8)


class WbOracle:
    """One whole-body OCP instance on the CPU oracle."""

    def __init__(self, model: dict):
        from wb_humanoid_mpc_b200 import abi

        self.model = model
        self.desc = abi.model_desc(model)
        self.nx, self.nu, self.nj = model["nx"], model["nu"], model["nj"]
        L = lib()
        L.orc_wb_create.restype = C.c_void_p
        L.orc_wb_total_mass.restype = C.c_double
        L.orc_wb_cost.restype = C.c_double
        L.orc_wb_cost_quad.restype = C.c_double
        self.h = C.c_void_p(L.orc_wb_create(C.byref(self.desc)))
        self.L = L
        self.n_nodes = 0

    def __del__(self):
        try:
            self.L.orc_wb_destroy(self.h)
        except Exception:
            pass

    def set_nodes(self, contact, swing, impact, arm_phase, xref):
        self.n_nodes = len(arm_phase)
        self._nodes = [np.ascontiguousarray(contact, dtype=np.uint8), F(swing), F(impact), F(arm_phase), F(xref)]
        c, s, i, a, x = self._nodes
        self.L.orc_wb_set_nodes(self.h, C.c_int(self.n_nodes), c.ctypes.data_as(u8p), _p(s), _p(i), _p(a), _p(x))

    def total_mass(self):
        return self.L.orc_wb_total_mass(self.h)

    def flow_map(self, x, u):
        out = np.zeros(self.nx)
        self.L.orc_wb_flow_map(self.h, _p(F(x)), _p(F(u)), _p(out))
        return out

    def base_accel_literal(self, x, u):
        out = np.zeros(6)
        self.L.orc_wb_base_accel_literal(self.h, _p(F(x)), _p(F(u)), _p(out))
        return out

    def joint_torques(self, x, u):
        tau, qddb = np.zeros(self.nj), np.zeros(6)
        self.L.orc_wb_joint_torques(self.h, _p(F(x)), _p(F(u)), _p(tau), _p(qddb))
        return tau, qddb

    def crba_nle(self, x):
        nv = 6 + self.nj
        M, nle = np.zeros((nv, nv)), np.zeros(nv)
        self.L.orc_wb_crba_nle(self.h, _p(F(x)), _p(M), _p(nle))
        return M, nle

    def rnea(self, q, v, a):
        tau = np.zeros(6 + self.nj)
        self.L.orc_wb_rnea(self.h, _p(F(q)), _p(F(v)), _p(F(a)), _p(tau))
        return tau

    def foot_state(self, x, u):
        out = np.zeros((2, 27))
        fp = np.zeros((self.desc.n_frames, 3))
        self.L.orc_wb_foot_state(self.h, _p(F(x)), _p(F(u)), _p(out), _p(fp))
        feet = []
        for c in range(2):
            o = out[c]
            feet.append(dict(pos=o[0:3], oriErr=o[3:6], vlin=o[6:9], vang=o[9:12], alin=o[12:15], aang=o[15:18], R=o[18:27].reshape(3, 3)))
        return feet, fp

    def flow_map_lin(self, x, u):
        f, A, B = np.zeros(self.nx), np.zeros((self.nx, self.nx)), np.zeros((self.nu, self.nx))
        self.L.orc_wb_flow_map_lin(self.h, _p(F(x)), _p(F(u)), _p(f), _p(A), _p(B))
        return f, A.T.copy(), B.T.copy()

    # -- centroidal flow map (oracle/cen_dynamics.hpp) ---------------------------------------------------------------------
    def centroidal_map(self, q):
        nv = 6 + self.nj
        Ag, com = np.zeros((6, nv)), np.zeros(3)
        self.L.orc_cen_centroidal_map(self.h, _p(F(q)), _p(Ag), _p(com))
        return Ag, com

    def cen_flow_map(self, x, u):
        xd = np.zeros(12 + self.nj)
        self.L.orc_cen_flow_map(self.h, _p(F(x)), _p(F(u)), _p(xd))
        return xd

    def cen_flow_map_lin(self, x, u):
        nx = nu = 12 + self.nj
        f, A, B = np.zeros(nx), np.zeros((nx, nx)), np.zeros((nu, nx))
        self.L.orc_cen_flow_map_lin(self.h, _p(F(x)), _p(F(u)), _p(f), _p(A), _p(B))
        return f, A.T.copy(), B.T.copy()

    def cost(self, k, x, u):
        return self.L.orc_wb_cost(self.h, C.c_int(k), _p(F(x)), _p(F(u)))

    def cost_quad(self, k, x, u):
        nx, nu = self.nx, self.nu
        Q, S, R, q, r = np.zeros((nx, nx)), np.zeros((nx, nu)), np.zeros((nu, nu)), np.zeros(nx), np.zeros(nu)
        f = self.L.orc_wb_cost_quad(self.h, C.c_int(k), _p(F(x)), _p(F(u)), _p(Q), _p(S), _p(R), _p(q), _p(r))
        return dict(f=f, Q=Q.T.copy(), S=S.T.copy(), R=R.T.copy(), q=q, r=r)

    def eq_constraint_lin(self, k, x, u):
        nx, nu = self.nx, self.nu
        g, Cm, Dm = np.zeros(14), np.zeros((nx, 14)), np.zeros((nu, 14))
        nc = self.L.orc_wb_eq_constraint_lin(self.h, C.c_int(k), _p(F(x)), _p(F(u)), _p(g), _p(Cm), _p(Dm), C.c_int(14))
        return g[:nc].copy(), Cm.T[:nc].copy(), Dm.T[:nc].copy()

    def eq_constraint(self, k, x, u):
        g = np.zeros(14)
        nc = self.L.orc_wb_eq_constraint(self.h, C.c_int(k), _p(F(x)), _p(F(u)), _p(g))
        return g[:nc].copy()

    def sqp(self, t_nodes, events, x0, x, u, settings, keep_raw=False):
        n = len(t_nodes)
        xs, us = F(x).copy(), F(u).copy()
        log = np.zeros((64, 16))
        nit = C.c_int(0)
        ev = np.ascontiguousarray(events, dtype=np.uint8)
        rc = self.L.orc_wb_sqp(self.h, C.c_int(n), _p(F(t_nodes)), ev.ctypes.data_as(u8p), _p(F(x0)), _p(xs), _p(us), C.byref(settings),
                               C.c_int(int(keep_raw)), _p(log), C.c_int(64), C.byref(nit))
        if rc != 0:
            raise RuntimeError("oracle SQP failed")
        dx, du = np.zeros((n, self.nx)), np.zeros((n - 1, self.nu))
        K = np.zeros((n - 1, self.nx, self.nu))
        nut = np.zeros(n - 1, dtype=np.int32)
        self.L.orc_wb_last_qp(self.h, _p(dx), _p(du), _p(K), _pi(nut))
        return dict(x=xs, u=us, log=log[: nit.value], dx=dx, du=du, K=np.swapaxes(K, 1, 2).copy(), nut=nut)

    def last_value_function(self, x_lin):
        n = len(x_lin)
        P, p = np.zeros((n, self.nx, self.nx)), np.zeros((n, self.nx))
        rc = self.L.orc_wb_last_value_function(self.h, _p(F(x_lin)), _p(P), _p(p))
        assert rc == 0
        return np.swapaxes(P, 1, 2).copy(), p

    def raw_per_node(self):
        nx, nu = self.nx, self.nu
        return 2 * nx * nx + 2 * nx * nu + nu * nu + 2 * nx + nu + 1 + 14 * (nx + nu + 1) + 1

    def last_raw_blocks(self, n_nodes):
        per = self.raw_per_node()
        out = np.zeros((n_nodes - 1, per))
        rc = self.L.orc_wb_last_raw_blocks(self.h, _p(out), C.c_longlong(per))
        assert rc == 0, rc
        return [unpack_raw_blocks(out[i], self.nx, self.nu) for i in range(n_nodes - 1)]


class CenOracle(WbOracle):
    """One centroidal OCP instance on the CPU oracle (oracle/cen_problem.hpp).  cost / cost_quad / eq_constraint(_lin) / sqp / last_* are the
    generic entry points of the base class; the whole-body-only queries are not available."""

    def __init__(self, model: dict):
        from wb_humanoid_mpc_b200 import abi

        assert model.get("kind") == "centroidal"
        self.model = model
        self.desc = abi.model_desc(model)
        self.cdesc = abi.cen_desc(model)
        self.nx, self.nu, self.nj = model["nx"], model["nu"], model["nj"]
        L = lib()
        L.orc_cen_create.restype = C.c_void_p
        L.orc_wb_total_mass.restype = C.c_double
        L.orc_wb_cost.restype = C.c_double
        L.orc_wb_cost_quad.restype = C.c_double
        self.h = C.c_void_p(L.orc_cen_create(C.byref(self.desc), C.byref(self.cdesc)))
        self.L = L
        self.n_nodes = 0

    def flow_map(self, x, u):
        return self.flow_map_lin(x, u)[0]

    def task_space(self, x, u):
        feet, torso, com, v = np.zeros((2, 12)), np.zeros(13), np.zeros(3), np.zeros(6 + self.nj)
        self.L.orc_cen_task_space(self.h, _p(F(x)), _p(F(u)), _p(feet), _p(torso), _p(com), _p(v))
        fs = [dict(pos=f[0:3], oriErr=f[3:6], vlin=f[6:9], vang=f[9:12]) for f in feet]
        return fs, dict(pos=torso[0:3], quat=torso[3:7], vlin=torso[7:10], vang=torso[10:13]), com, v

    def residuals(self, k, x, u, jacobian=True):
        nz = self.nx + self.nu
        r, J = np.zeros(64), np.zeros((64, nz))
        n = self.L.orc_cen_residuals(self.h, C.c_int(k), _p(F(x)), _p(F(u)), _p(r), _p(J) if jacobian else None)
        return r[:n].copy(), J[:n].copy()


def unpack_raw_blocks(p, nx, nu):
    """Layout shared by the oracle (orc_wb_last_raw_blocks) and the CUDA path (b200sqp_download_stage_blocks which=0)."""
    o = 0

    def take(n):
        nonlocal o
        v = p[o:o + n]
        o += n
        return v

    A = take(nx * nx).reshape(nx, nx).T
    B = take(nx * nu).reshape(nu, nx).T
    b = take(nx)
    Q = take(nx * nx).reshape(nx, nx).T
    S = take(nu * nx).reshape(nx, nu).T
    R = take(nu * nu).reshape(nu, nu).T
    q = take(nx)
    r = take(nu)
    c = take(1)[0]
    Cm = take(14 * nx).reshape(nx, 14).T
    Dm = take(14 * nu).reshape(nu, 14).T
    e = take(14)
    nc = int(round(take(1)[0]))
    return dict(A=A, B=B, b=b, Q=Q, S=S, R=R, q=q, r=r, c=c, C=Cm[:nc], D=Dm[:nc], e=e[:nc], nc=nc)
