"""CPU oracle of the centroidal OCP (oracle/cen_problem.hpp): the reference has no stored numbers for the G1 terms, so the
restatement is pinned by finite differences of every Jacobian, by kinematic identities and by the behaviour of the SQP on BASELINE
configs[0] (N = 20) -- the same strategy as tests/test_oracle_wb.py."""
import numpy as np
import pytest

import oracle_lib as ol
from wb_humanoid_mpc_b200 import abi, references as R
from wb_humanoid_mpc_b200.model_loader import load_packaged_model


@pytest.fixture(scope="module")
def model():
    return load_packaged_model("g1_centroidal")


@pytest.fixture(scope="module")
def orc(model):
    return ol.CenOracle(model)


def perturbed_state(model, rng, scale=1.0):
    x = np.array(model["x_init"])
    x[:6] += scale * rng.uniform(-0.2, 0.2, 6)
    x[6:9] += scale * rng.uniform(-0.02, 0.02, 3)
    x[9:12] += scale * rng.uniform(-0.05, 0.05, 3)
    x[12:] += scale * rng.uniform(-0.1, 0.1, model["nj"])
    return x


def random_input(model, rng, contacts):
    u = R.weight_compensating_input(model, contacts) + rng.uniform(-5, 5, model["nu"])
    u[12:] = rng.uniform(-0.5, 0.5, model["nj"])
    return u


def walk_instance(model, horizon=0.4, cmd=(0.3, 0.1, 0.7925, 0.2), x0=None, **kw):
    x0 = np.array(model["x_init"]) if x0 is None else x0
    return R.build_instance(model, x0, gait="walk", cmd=list(cmd), horizon=horizon, **kw)


def set_nodes(orc, I):
    orc.set_nodes(I["contact_flags"], I["swing_ref"], I["impact_factor"], I["arm_phase"], I["x_ref"])


def fd(fun, z, eps=1e-6):
    f0 = fun(z)
    J = np.zeros((len(f0), len(z)))
    for i in range(len(z)):
        zp, zm = z.copy(), z.copy()
        zp[i] += eps
        zm[i] -= eps
        J[:, i] = (fun(zp) - fun(zm)) / (2 * eps)
    return J


def test_model_dimensions(model):
    assert model["nx"] == 35 and model["nu"] == 35 and model["kind"] == "centroidal"
    assert model["sqp"]["dt"] == 0.02 and model["sqp"]["timeHorizon"] == 1.2
    assert model["task_space_cost"]["link"] == "mid360_link"


def test_frame_velocity_is_jacobian_times_generalized_velocity(model, orc):
    """v_frame = d/dt p_frame(q(t)) along qdot = getPinocchioJointVelocity(x, u)"""
    rng = np.random.default_rng(0)
    x, u = perturbed_state(model, rng), random_input(model, rng, [1, 1])
    feet, torso, com, v = orc.task_space(x, u)
    eps = 1e-6
    xp, xm = x.copy(), x.copy()
    xp[6:] += eps * v
    xm[6:] -= eps * v
    fp, tp, cp, _ = orc.task_space(xp, u)
    fm, tm, cm, _ = orc.task_space(xm, u)
    for c in range(2):
        assert np.allclose((fp[c]["pos"] - fm[c]["pos"]) / (2 * eps), feet[c]["vlin"], atol=1e-7)
    assert np.allclose((tp["pos"] - tm["pos"]) / (2 * eps), torso["vlin"], atol=1e-7)
    # centre of mass velocity = normalized linear momentum (definition of the centroidal state)
    assert np.allclose((cp - cm) / (2 * eps), x[:3], atol=1e-7)
    # quaternion of the link: unit norm, w > 0 near upright, and consistent with the angular velocity: qdot = 1/2 [w] (x) q
    assert abs(np.linalg.norm(torso["quat"]) - 1) < 1e-12 and torso["quat"][3] > 0.9
    qd = (tp["quat"] - tm["quat"]) / (2 * eps)
    w, q = torso["vang"], torso["quat"]
    expect = 0.5 * np.concatenate([q[3] * w + np.cross(w, q[:3]), [-w @ q[:3]]])
    assert np.allclose(qd, expect, atol=1e-7)


def test_equality_constraints_shapes_and_values(model, orc):
    I = walk_instance(model)
    set_nodes(orc, I)
    x = np.array(model["x_init"])
    k = 1
    assert list(I["contact_flags"][k]) == [1, 0]
    u = R.weight_compensating_input(model, I["contact_flags"][k])
    g = orc.eq_constraint(k, x, u)
    assert len(g) == 6 + 7
    feet, *_ = orc.task_space(x, u)
    # stance foot at rest: twist rows vanish except for the height and orientation feedback
    assert np.allclose(g[:2], 0) and np.isclose(g[2], 5.0 * feet[0]["pos"][2]) and np.allclose(g[3:6], 20.0 * feet[0]["oriErr"])
    # swing foot: zero wrench rows, then the normal velocity row  v_z - vref + 5 (p_z - pref)
    assert np.allclose(g[6:12], u[6:12])
    sw = I["swing_ref"][k, 1]
    assert np.isclose(g[12], feet[1]["vlin"][2] - sw[1] + 5.0 * (feet[1]["pos"][2] - sw[0]))


@pytest.mark.parametrize("k", [1, 12])
def test_constraint_jacobians_match_finite_differences(model, orc, k):
    I = walk_instance(model, horizon=0.6)
    set_nodes(orc, I)
    rng = np.random.default_rng(k)
    x, u = perturbed_state(model, rng), random_input(model, rng, I["contact_flags"][k])
    g, Cm, Dm = orc.eq_constraint_lin(k, x, u)
    nx = model["nx"]
    J = fd(lambda z: orc.eq_constraint(k, z[:nx], z[nx:]), np.concatenate([x, u]))
    assert np.allclose(Cm, J[:, :nx], atol=2e-6) and np.allclose(Dm, J[:, nx:], atol=2e-6)
    assert np.allclose(g, orc.eq_constraint(k, x, u))


@pytest.mark.parametrize("k", [1, 12])
def test_residual_jacobians_match_finite_differences(model, orc, k):
    m2 = dict(model)
    m2["icp_weight"] = 3.0  # exercise the ICP term as well (weight 0 in the shipped file)
    o2 = ol.CenOracle(m2)
    I = walk_instance(model, horizon=0.6)
    set_nodes(o2, I)
    rng = np.random.default_rng(10 + k)
    x, u = perturbed_state(model, rng), random_input(model, rng, I["contact_flags"][k])
    r, J = o2.residuals(k, x, u)
    ns = int(I["contact_flags"][k].sum())
    assert len(r) == 12 + 2 + 24 + 6 * ns
    nx = model["nx"]
    Jfd = fd(lambda z: o2.residuals(k, z[:nx], z[nx:], jacobian=False)[0], np.concatenate([x, u]))
    assert np.allclose(J, Jfd, atol=5e-6)
    assert np.abs(r[12:14]).max() > 1e-3  # ICP rows are live with a non-zero weight


def test_task_space_residual_vanishes_on_the_reference(model, orc):
    I = walk_instance(model, cmd=(0.0, 0.0, 0.7925, 0.0))
    set_nodes(orc, I)
    k = 3
    r, _ = orc.residuals(k, I["x_ref"][k], np.zeros(model["nu"]))
    assert np.allclose(r[:12], 0, atol=1e-12)
    assert np.allclose(r[12:14], 0)  # ICP weight 0


@pytest.mark.parametrize("k", [1, 12])
def test_cost_gradient_and_gauss_newton_hessian(model, orc, k):
    I = walk_instance(model, horizon=0.6)
    set_nodes(orc, I)
    rng = np.random.default_rng(20 + k)
    x, u = perturbed_state(model, rng), random_input(model, rng, I["contact_flags"][k])
    c = orc.cost_quad(k, x, u)
    assert np.isclose(c["f"], orc.cost(k, x, u), rtol=1e-13)
    nx = model["nx"]
    z = np.concatenate([x, u])

    # the quadratic tracking cost treats the arm-swing reference (a function of the current yaw) as a constant: remove that dependence
    def cost_frozen(zz):
        return np.array([orc.cost(k, zz[:nx], zz[nx:])])

    g = fd(cost_frozen, z, eps=1e-6)[0]
    grad = np.concatenate([c["q"], c["r"]])
    yaw_leak = np.abs(g - grad)
    yaw_leak[9] = 0.0
    assert yaw_leak.max() < 2e-4 * max(1.0, np.abs(grad).max())
    H = np.block([[c["Q"], c["S"].T], [c["S"], c["R"]]])  # S = dfdux (nu x nx)
    assert np.allclose(H, H.T, atol=1e-9)
    ev = np.linalg.eigvalsh(H)
    assert ev.min() > -1e-6 * ev.max()  # Gauss-Newton + barrier Hessians: positive semi-definite up to the friction-cone curvature shift


def test_external_torque_rows_follow_the_wrench(model, orc):
    """J_ee' W: a pure vertical force through the ankle-roll axis origin produces no ankle torque but loads hip pitch and knee"""
    I = walk_instance(model)
    set_nodes(orc, I)
    k = 1
    x = np.array(model["x_init"])
    u = R.weight_compensating_input(model, I["contact_flags"][k])
    r0, J = orc.residuals(k, x, u)
    assert len(r0) == 12 + 2 + 24 + 6
    tq = r0[12 + 2 + 12:12 + 2 + 12 + 6]
    mid = 1.0 - I["impact_factor"][k, 1]
    w = np.sqrt(np.array(model["leg_torque_cost"][0]["weights"]))
    # residual = sqrt(w) * mid * tau: tau linear in the wrench -> doubling the force doubles the rows
    r1, _ = orc.residuals(k, x, 2 * u)
    assert np.allclose(r1[26:32], 2 * tq, atol=1e-12)
    if mid > 0:
        tau = tq / (w * mid)
        assert abs(tau[3]) > abs(tau[5])  # knee carries more than the ankle roll


def test_sqp_config0_converges(model, orc):
    """BASELINE configs[0]: N = 20, one instance; the SQP drives the constraint violation down and keeps the merit bounded"""
    I = walk_instance(model, horizon=0.4)
    set_nodes(orc, I)
    st = abi.default_settings(model, sqp_iteration=8)
    out = orc.sqp(I["t_nodes"], I["node_event"], I["x0"], I["x_init"], I["u_init"], st)
    log = out["log"]
    assert len(log) >= 2
    viol0 = log[0][2] + log[0][3]
    viol1 = log[-1][6] + log[-1][7]
    assert viol1 < 1e-3 * max(viol0, 1e-6) or viol1 < 1e-9
    assert np.all(np.isfinite(out["x"])) and np.all(np.isfinite(out["u"]))
    # equality constraints hold on the solution
    for k in range(len(I["t_nodes"]) - 1):
        if I["node_event"][k] == 1:
            continue
        assert np.abs(orc.eq_constraint(k, out["x"][k], out["u"][k])).max() < 5e-2


# ---- SingleRigidBodyDynamics model type (centroidalModelType 1) ----------------------------------------------------------------------------
def srbd_model(model):
    m = dict(model)
    m["centroidalModelType"] = 1
    return m


def test_srbd_equals_full_model_at_nominal_joints(model, orc):
    """The SRBD model freezes the centroidal inertia and the com offset at the nominal joint angles: there (any base pose, zero joint
    velocities) both model types give the same flow map.  Pins model_loader.srbd_nominal against the oracle's centroidal momentum matrix."""
    o1 = ol.CenOracle(srbd_model(model))
    rng = np.random.default_rng(30)
    x = np.array(model["x_init"], float)
    x[12:] = model["reference"]["defaultJointState"]
    x[:6] = rng.uniform(-0.3, 0.3, 6)
    x[6:12] += rng.uniform(-0.3, 0.3, 6)
    u = random_input(model, rng, [1, 1])
    u[12:] = 0.0
    assert np.allclose(o1.flow_map(x, u), orc.flow_map(x, u), atol=1e-12)
    u[12:] = rng.uniform(-0.5, 0.5, model["nj"])   # joint velocities move the full model's base (Aj qdot_j), not the SRBD one
    f1, f0 = o1.flow_map(x, u), orc.flow_map(x, u)
    assert np.allclose(f1[:6], f0[:6], atol=1e-12) and np.abs(f1[6:12] - f0[6:12]).max() > 1e-3


def test_srbd_flow_map_jacobians_and_base_velocity(model):
    from wb_humanoid_mpc_b200 import centroidal

    m1 = srbd_model(model)
    o1 = ol.CenOracle(m1)
    rng = np.random.default_rng(31)
    x, u = perturbed_state(model, rng), random_input(model, rng, [1, 0])
    f, A, B = o1.flow_map_lin(x, u)
    nx = model["nx"]
    J = fd(lambda z: o1.flow_map(z[:nx], z[nx:]), np.concatenate([x, u]))
    assert np.allclose(A, J[:, :nx], atol=2e-6) and np.allclose(B, J[:, nx:], atol=2e-6)
    # closed-form Ab^-1 hbar of the target-trajectory rule == flow(x, 0)[6:12] / m
    bv = centroidal.base_velocity(m1, x)
    assert np.allclose(bv, o1.flow_map(x, np.zeros(model["nu"]))[6:12] / sum(model["mass"]), atol=1e-12)
