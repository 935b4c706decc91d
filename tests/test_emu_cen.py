"""Device math of the centroidal path without a GPU: the __host__ __device__ phase functions of cen_lq_kernel / cen_proj_kernel executed on
the host (tests/emu, work items in reverse order) against the oracle's LQ blocks (oracle/cen_problem.hpp).  The real launches are checked
by tests/test_gpu_cen_ocp.py."""
import numpy as np
import pytest

import emu_lib as emu
import oracle_lib as orc
from test_oracle_cen_ocp import perturbed_state, random_input
from wb_humanoid_mpc_b200 import abi, model_loader


@pytest.fixture(scope="module", params=[0, 1], ids=["full", "srbd"])
def model(request):
    m = dict(model_loader.load_packaged_model("g1_centroidal"))
    m["icp_weight"] = 2.0   # exercise the ICP rows too (0 in the shipped task.info)
    m["centroidalModelType"] = request.param   # 0 FullCentroidalDynamics (shipped), 1 SingleRigidBodyDynamics
    return m


@pytest.fixture(scope="module")
def cen(model):
    return orc.CenOracle(model)


def rel(a, b):
    return np.max(np.abs(a - b)) / max(1e-12, np.max(np.abs(b))) if b.size else 0.0


@pytest.mark.parametrize("contact", [(1, 1), (1, 0), (0, 1), (0, 0)])
def test_cen_node_blocks_and_projection(model, cen, contact):
    rng = np.random.default_rng(sum(contact) + 7)
    x, u = perturbed_state(model, rng, 0.5), random_input(model, rng, contact)
    xn = x + rng.uniform(-0.01, 0.01, 35)
    xref = np.array(model["x_init"])
    xref[0:2] = [0.4, 0.1]
    xref[9] = 0.05
    swing = np.array([[0.03, 0.2, -1.0], [0.05, -0.1, 0.5]])
    impact, arm, dt = [0.7, 0.4], 0.3, 0.02
    cen.set_nodes(np.array([contact, contact], dtype=np.uint8), np.stack([swing, swing]), np.array([impact, impact]), np.array([arm, arm]),
                  np.stack([xref, xref]))
    st = abi.default_settings(model, sqp_iteration=1)
    cen.sqp(np.array([0.0, dt]), np.array([0, 0], dtype=np.uint8), x, np.stack([x, xn]), u[None], st, keep_raw=True)
    raw = cen.last_raw_blocks(2)[0]
    e = emu.cen_node(cen.desc, cen.cdesc, x, u, xn, xref, dt, contact, swing, impact, arm)
    assert e["raw"]["nc"] == raw["nc"] and e["nut"] == 35 - raw["nc"]
    for k in ["A", "B", "b", "Q", "S", "R", "q", "r", "C", "D", "e"]:
        assert rel(e["raw"][k], raw[k]) < 1e-9, k
    assert abs(e["raw"]["c"] - raw["c"]) < 1e-10 * max(1.0, abs(raw["c"]))
    r = e["raw"]
    Px, Pu, u0 = e["Px"][:, :35], e["Pu"], e["u0"]
    assert np.abs(e["Px"][:, 35:]).max() == 0.0
    assert np.abs(r["D"] @ Pu).max() < 1e-10 and np.abs(r["D"] @ Px + r["C"]).max() < 1e-9 and np.abs(r["D"] @ u0 + r["e"]).max() < 1e-9
    pr = orc.change_of_input_variables(r["A"], r["B"], r["b"], r["Q"], r["S"], r["R"], r["q"], r["r"], r["c"], Pu, Px, u0)
    # live block of the padded record; the dummy states are inert: A = 1 on their diagonal, zero everywhere else
    assert rel(e["A"][:35, :35], pr["A"]) < 1e-10 and np.array_equal(e["A"][35:, 35:], np.eye(23)) and not e["A"][:35, 35:].any() and not e["A"][35:, :35].any()
    assert rel(e["B"][:35], pr["B"]) < 1e-10 and not e["B"][35:].any()
    assert rel(e["b"][:35], pr["b"]) < 1e-10 and not e["b"][35:].any()
    assert rel(e["Q"][:35, :35], pr["Q"]) < 1e-10 and not e["Q"][35:].any() and not e["Q"][:, 35:].any()
    assert rel(e["S"][:, :35], pr["S"]) < 1e-10 and not e["S"][:, 35:].any()
    assert rel(e["R"], pr["R"]) < 1e-10 and rel(e["q"][:35], pr["q"]) < 1e-10 and rel(e["r"], pr["r"]) < 1e-10
