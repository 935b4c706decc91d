"""include/ocs2_sqp/B200SqpSolver.h -- the ocs2::SolverBase drop-in a wb_humanoid_mpc workspace puts behind SqpMpc --
compiled with g++ against the stand-in ocs2 headers of tests/stubs/ and driven like SqpMpc::calculateController drives ocs2::SqpSolver
(tests/stubs/shim_test.cpp).  CPU: it compiles, links against libb200sqp.so and refuses to run without a device (no CPU fallback).  GPU: two MPC
cycles (cold, then warm start from its own PrimalSolution); the first solution equals the batched C++ host layer's on the same instance."""
import subprocess
from pathlib import Path

import pytest

from wb_humanoid_mpc_b200 import lib

ROOT = Path(__file__).resolve().parents[1]
BUILD = ROOT / "tests" / "stubs" / "_build"


def build_shim_test():
    lib.lib()   # make sure libb200sqp.so exists
    BUILD.mkdir(exist_ok=True)
    exe = BUILD / "shim_test"
    srcs = [ROOT / "tests" / "stubs" / "shim_test.cpp", ROOT / "tests" / "stubs" / "ocs2_stub.hpp", ROOT / "include" / "ocs2_sqp" / "B200SqpSolver.h",
            ROOT / "include" / "b200sqp.h"] + list((ROOT / "wb_humanoid_mpc_b200" / "host").glob("*.hpp"))
    if not exe.exists() or any(s.stat().st_mtime > exe.stat().st_mtime for s in srcs):
        pkg = ROOT / "wb_humanoid_mpc_b200"
        subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", f"-I{ROOT / 'include'}", f"-I{ROOT / 'tests' / 'stubs'}", "-o", str(exe), str(srcs[0]), f"-L{pkg}", "-lb200sqp",
                        "-pthread", f"-Wl,-rpath,{pkg}"], check=True, cwd=ROOT)
    return exe


def run_shim():
    exe = build_shim_test()
    return subprocess.run([str(exe), str(ROOT / "wb_humanoid_mpc_b200" / "data" / "g1_wb_model.txt")], capture_output=True, text=True, cwd=ROOT, timeout=300)


def test_shim_compiles_against_stub_ocs2_and_fails_loudly_without_a_device():
    import torch

    out = run_shim()
    if torch.cuda.is_available():
        assert out.returncode == 0, out.stdout + out.stderr
    else:
        assert out.returncode == 3 and "no CUDA device" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_shim_two_mpc_cycles_match_the_host_layer():
    out = run_shim()
    assert out.returncode == 0, out.stdout + out.stderr
    assert "SHIM_OK" in out.stdout
