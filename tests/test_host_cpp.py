"""The C++ host layer (wb_humanoid_mpc_b200/host/*.hpp) against the independent Python restatement (references.py): every per-node array of
b200sqp_upload_instances must agree for cold and warm starts, all gaits, shifted gait phases.  CPU only (no compute calls)."""
import numpy as np
import pytest

from wb_humanoid_mpc_b200 import host_lib, model_loader, references


@pytest.fixture(scope="module")
def model():
    return model_loader.load_packaged_model()


@pytest.fixture(scope="module")
def hmodel():
    m = host_lib.HostModel()
    yield m
    m.close()


def test_flat_model_file_matches_json(model, hmodel):
    from wb_humanoid_mpc_b200 import abi

    d_py = abi.model_desc(model)
    d_cc, st = hmodel.desc_and_settings()
    assert bytes(d_py) == bytes(d_cc), "b200sqp_model_desc read by the C++ host differs from the Python mirror"
    sp = abi.default_settings(model)
    assert bytes(sp) == bytes(st)
    assert (hmodel.nx, hmodel.nu, hmodel.dt, hmodel.horizon) == (model["nx"], model["nu"], model["sqp"]["dt"], model["sqp"]["timeHorizon"])


KEYS = ["t_nodes", "node_event", "contact_flags", "swing_ref", "impact_factor", "arm_phase", "x_ref", "x_init", "u_init"]


def x0_of(model, rng):
    x0 = np.array(model["x_init"], float)
    x0[:6] += rng.uniform(-0.05, 0.05, 6)
    x0[29:35] += rng.uniform(-0.3, 0.3, 6)
    return x0


@pytest.mark.parametrize("gait,t0,horizon,start", [("stance", 0.0, 1.1, None), ("walk", 0.0, 3.5, None), ("walk", 0.37, 2.0, 0.1),
                                                   ("slow_walk", 1.25, 3.5, 0.5), ("trot", 0.0, 1.1, None), ("left_leg", 0.2, 1.5, 0.2)])
def test_cold_start_instance_matches_python(model, hmodel, gait, t0, horizon, start):
    rng = np.random.default_rng(7)
    x0 = x0_of(model, rng)
    cmd = [0.6, -0.1, model["reference"]["defaultBaseHeight"], 0.3]
    py = references.build_instance(model, x0, t0=t0, horizon=horizon, gait=gait, gait_start=start, cmd=cmd)
    cc = hmodel.build_instance(x0, t0=t0, horizon=horizon, gait=gait, gait_start=start, cmd=cmd)
    for k in KEYS:
        a, b = np.asarray(py[k]), np.asarray(cc[k])
        assert a.shape == b.shape, (k, a.shape, b.shape)
        assert np.allclose(a, b, rtol=0, atol=1e-13), (k, np.abs(a.astype(float) - b.astype(float)).max())


def test_warm_start_instance_matches_python(model, hmodel):
    rng = np.random.default_rng(9)
    x0 = x0_of(model, rng)
    first = references.build_instance(model, x0, t0=0.0, horizon=1.1, gait="walk")
    n = len(first["t_nodes"])
    # a made-up previous solution on the first grid (values are irrelevant to the interpolation logic)
    xs, us = rng.normal(size=(n, 58)), rng.normal(size=(n - 1, 35))
    prev = references.to_primal_solution(first["t_nodes"], first["node_event"], xs, us)
    x1 = x0 + 0.01
    py = references.build_instance(model, x1, t0=0.105, horizon=1.1, gait="walk", gait_start=0.0, previous=prev)
    cc = hmodel.build_instance(x1, t0=0.105, horizon=1.1, gait="walk", gait_start=0.0, previous=prev)
    for k in KEYS:
        assert np.allclose(np.asarray(py[k], float), np.asarray(cc[k], float), rtol=0, atol=1e-13), k
    # the overlap really is interpolated, the tail really is the weight-compensating initializer
    assert not np.allclose(cc["u_init"][0], cc["u_init"][-1])


def test_errors_are_reported(hmodel):
    with pytest.raises(RuntimeError, match="unknown gait"):
        hmodel.build_instance(np.zeros(58), gait="moonwalk")
    with pytest.raises(RuntimeError, match="not found"):
        host_lib.HostModel("/nonexistent/model.txt")
