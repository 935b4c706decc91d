"""ctypes bindings of the CPU development harness (tests/emu/libwbemu.so): the CUDA kernels' __host__ __device__ phase functions
executed on the host.  Test infrastructure only -- never linked into or called by the product library."""
from __future__ import annotations

import ctypes as C
import shutil
import subprocess
from pathlib import Path

import numpy as np

from oracle_lib import F, _p, unpack_raw_blocks

ROOT = Path(__file__).resolve().parents[1]
EMU = ROOT / "tests" / "emu"
CSRC = ROOT / "wb_humanoid_mpc_b200" / "csrc"
_lib = None


def build(force=False):
    so = EMU / "libwbemu.so"
    srcs = list(EMU.glob("*.cu")) + list(EMU.glob("*.inc")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.inc"))
    if force or not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        subprocess.run([nvcc, "-O2", "-std=c++17", "-DEMU_WITH_LQ", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-Wno-deprecated-gpu-targets",
                        "-diag-suppress", "177", "-o", str(so), str(EMU / "wb_emu.cu")], check=True)
    return so


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
    return _lib


class EmuNodeIn(C.Structure):
    _fields_ = [("x", C.POINTER(C.c_double)), ("u", C.POINTER(C.c_double)), ("xnext", C.POINTER(C.c_double)), ("xref", C.POINTER(C.c_double)),
                ("dt", C.c_double), ("contact", C.c_int * 2), ("swing", C.c_double * 6), ("impact", C.c_double * 2), ("armPhase", C.c_double)]


def dyn(desc, x, u, deriv=True):
    xd, G = np.zeros(58), np.zeros((93, 6))
    rc = lib().emu_dyn(C.byref(desc), _p(F(x)), _p(F(u)), _p(xd), _p(G), C.c_int(int(deriv)))
    assert rc == 0
    return xd, G.T.copy()


def lq_node(desc, x, u, xnext, xref, dt, contact, swing, impact, arm_phase, mode=0):
    """mode 0: one-pass K1b schedule; 1 / 2: the CUDA schedules (lu + part 1 + part 2; 2 = Q accumulated in place)."""
    NX, NU, NT = 58, 35, 23
    xs, us, xn, xr = F(x), F(u), F(xnext), F(xref)
    nin = EmuNodeIn(_p(xs), _p(us), _p(xn), _p(xr), dt, (C.c_int * 2)(*[int(c) for c in contact]), (C.c_double * 6)(*np.asarray(swing, float).reshape(6)),
                    (C.c_double * 2)(*impact), arm_phase)
    A, Bt, b = np.zeros((NX, NX)), np.zeros((NT, NX)), np.zeros(NX)
    Q, St, Rt, q, rt = np.zeros((NX, NX)), np.zeros((NX, NT)), np.zeros((NT, NT)), np.zeros(NX), np.zeros(NT)
    Pu, Px, u0 = np.zeros((NT, NU)), np.zeros((NX, NU)), np.zeros(NU)
    nut = C.c_int(0)
    perf = np.zeros(4)
    per = 2 * NX * NX + 2 * NX * NU + NU * NU + 2 * NX + NU + 1 + 14 * (NX + NU + 1) + 1
    raw = np.zeros(per)
    rc = lib().emu_lq_node_mode(C.byref(desc), C.byref(nin), _p(A), _p(Bt), _p(b), _p(Q), _p(St), _p(Rt), _p(q), _p(rt), _p(Pu), _p(Px), _p(u0),
                                C.byref(nut), _p(perf), _p(raw), C.c_int(mode))
    assert rc == 0
    n = nut.value
    return dict(A=A.T.copy(), B=Bt.T[:, :n].copy(), b=b, Q=Q.T.copy(), S=St.T[:n].copy(), R=Rt.T[:n, :n].copy(), q=q, r=rt[:n].copy(),
                Pu=Pu.T[:, :n].copy(), Px=Px.T.copy(), u0=u0, nut=n, perf=perf, raw=unpack_raw_blocks(raw, NX, NU))


def cen_node(desc, cdesc, x, u, xnext, xref, dt, contact, swing, impact, arm_phase):
    """One intermediate node of the centroidal path: cen_lq_kernel + cen_proj_kernel phase schedules on the host (padded QP record)."""
    NX, NU, NT, CX = 58, 35, 23, 35
    pad = lambda v: np.concatenate([F(v), np.zeros(NX - CX)])
    xs, us, xn, xr = pad(x), F(u), pad(xnext), pad(xref)
    nin = EmuNodeIn(_p(xs), _p(us), _p(xn), _p(xr), dt, (C.c_int * 2)(*[int(c) for c in contact]), (C.c_double * 6)(*np.asarray(swing, float).reshape(6)),
                    (C.c_double * 2)(*impact), arm_phase)
    A, Bt, b = np.zeros((NX, NX)), np.zeros((NT, NX)), np.zeros(NX)
    Q, St, Rt, q, rt = np.zeros((NX, NX)), np.zeros((NX, NT)), np.zeros((NT, NT)), np.zeros(NX), np.zeros(NT)
    Pu, Px, u0 = np.zeros((NT, NU)), np.zeros((NX, NU)), np.zeros(NU)
    nut = C.c_int(0)
    perf = np.zeros(4)
    per = 2 * CX * CX + 2 * CX * NU + NU * NU + 2 * CX + NU + 1 + 14 * (CX + NU + 1) + 1
    raw = np.zeros(per)
    rc = lib().emu_cen_node(C.byref(desc), C.byref(cdesc), C.byref(nin), _p(raw), _p(A), _p(Bt), _p(b), _p(Q), _p(St), _p(Rt), _p(q), _p(rt),
                            _p(Pu), _p(Px), _p(u0), C.byref(nut), _p(perf))
    assert rc == 0
    n = nut.value
    return dict(A=A.T.copy(), B=Bt.T[:, :n].copy(), b=b, Q=Q.T.copy(), S=St.T[:n].copy(), R=Rt.T[:n, :n].copy(), q=q, r=rt[:n].copy(),
                Pu=Pu.T[:, :n].copy(), Px=Px.T.copy(), u0=u0, nut=n, perf=perf, raw=unpack_raw_blocks(raw, CX, NU))


def joint_torques(desc, x, u):
    tau, qddb = np.zeros(23), np.zeros(6)
    rc = lib().emu_joint_torques(C.byref(desc), _p(F(x)), _p(F(u)), _p(tau), _p(qddb))
    assert rc == 0
    return tau, qddb


def build_instances(model, t0, horizon, x0, gaits, gait_start, cmd, previous=None, cap=None):
    """the device-side instance builder (wb_builder.cuh) on the CPU harness -> dict in the layout of solver.stack_instances, or a negative error code.
    previous = dict(t [B, n], event [B, n], x [B, n, nx], u [B, n-1, nu]) enables the warm start."""
    from wb_humanoid_mpc_b200 import abi

    desc, names = abi.builder_desc(model)
    x0 = F(x0)
    B, nx, nu = x0.shape[0], model["nx"], model["nu"]
    gid = np.ascontiguousarray([names.index(g) for g in gaits], dtype=np.int32)
    gs, cm = F(gait_start), F(cmd)
    cap = cap or int(horizon / desc.dt) + 2 + 96
    out = dict(x_init=np.zeros((B, cap, nx)), u_init=np.zeros((B, cap, nu)), t_nodes=np.zeros((B, cap)), node_event=np.zeros((B, cap), dtype=np.uint8),
               contact_flags=np.zeros((B, cap, 2), dtype=np.uint8), swing_ref=np.zeros((B, cap, 2, 3)), impact_factor=np.zeros((B, cap, 2)),
               arm_phase=np.zeros((B, cap)), x_ref=np.zeros((B, cap, nx)))
    u8p = C.POINTER(C.c_uint8)
    ip = C.POINTER(C.c_int32)
    if previous is not None:
        pT, pE, pX, pU = F(previous["t"]), np.ascontiguousarray(previous["event"], dtype=np.uint8), F(previous["x"]), F(previous["u"])
        prevN = pT.shape[1] - 1
        pa = (_p(pT), pE.ctypes.data_as(u8p), _p(pX), _p(pU))
    else:
        prevN, pa = 0, (None, None, None, None)
    L = lib()
    L.emu_build_instances.restype = C.c_int
    n = L.emu_build_instances(C.byref(desc), C.c_int(B), C.c_double(t0), C.c_double(horizon), _p(x0), gid.ctypes.data_as(ip), _p(gs), _p(cm),
                              C.c_int(int(previous is not None)), C.c_int(prevN), *pa, C.c_int(cap), _p(out["x_init"]), _p(out["u_init"]), _p(out["t_nodes"]),
                              out["node_event"].ctypes.data_as(u8p), out["contact_flags"].ctypes.data_as(u8p), _p(out["swing_ref"]), _p(out["impact_factor"]),
                              _p(out["arm_phase"]), _p(out["x_ref"]))
    if n < 0:
        return n
    # the harness writes with the strides of the actual node count
    res = {}
    for k, v in out.items():
        per = int(np.prod(v.shape[2:])) if v.ndim > 2 else 1
        rows = n - 1 if k == "u_init" else n
        flat = v.reshape(-1)[: B * rows * per]
        res[k] = flat.reshape((B, rows) + v.shape[2:]).copy()
    res["x0"] = x0
    return res
