// ORACLE (test infrastructure only) -- CPU restatement of the Unitree G1 whole-body OCP terms.
//
// x = [q(6+nj); qdot(6+nj)], u = [W_left(6); W_right(6); qddot_joints(nj)]
//   (humanoid_nmpc/humanoid_wb_mpc/include/humanoid_wb_mpc/common/WBAccelMpcRobotModel.h:76-241)
// Terms and their reference sources (paths relative to /root/reference/humanoid_nmpc/):
//   dynamics            humanoid_wb_mpc/src/dynamics/DynamicsHelperFunctions.cpp:51-134 (computeBaseAcceleration, computeStateDerivative)
//                       humanoid_common_mpc/src/pinocchio_model/DynamicsHelperFunctions.cpp:196-218 (block-diagonal M_bb inverse)
//   quadratic tracking  humanoid_common_mpc/src/cost/StateInputQuadraticCost.cpp:67-78, reference_manager/SwitchedModelReferenceManager.cpp:110-135,
//                       include/humanoid_common_mpc/pinocchio_model/DynamicsHelperFunctions.h:178-193 (weightCompensatingInput)
//   foot GN cost        humanoid_wb_mpc/src/cost/EndEffectorDynamicsFootCost.cpp:91-152
//   friction cone       humanoid_common_mpc/src/constraint/FrictionForceConeConstraint.cpp:80-224
//   contact moment XY   humanoid_common_mpc/src/constraint/ContactMomentXYConstraintCppAd.cpp:86-104
//   joint limits        humanoid_common_mpc/src/constraint/JointLimitsSoftConstraint.cpp:69-100
//   foot collision      humanoid_common_mpc/src/constraint/FootCollisionConstraint.cpp:80-144
//   zero wrench         humanoid_common_mpc/src/constraint/ZeroWrenchConstraint.cpp:59-84
//   stance-foot accel.  humanoid_wb_mpc/src/constraint/EndEffectorDynamicsAccelerationsConstraint.cpp:84-146, WBMpcInterface.cpp:205-229
//   swing-foot z        humanoid_wb_mpc/src/constraint/EndEffectorDynamicsLinearAccConstraint.cpp:77-127, WBMpcPreComputation.cpp:68-113
//   penalties           lib/ocs2_ros2/ocs2_core/src/penalties/penalties/RelaxedBarrierPenalty.cpp:37-66, PieceWisePolynomialBarrierPenalty.cpp:37-75,
//                       MultidimensionalPenalty.cpp:143-226
//   OCP wiring / order  humanoid_wb_mpc/src/WBMpcInterface.cpp:131-199
#pragma once
#include <cstdint>
#include <memory>

#include "rbd.hpp"
#include "sqp.hpp"

namespace orc {

struct WbParams {
  RobotModel model;
  int nj = 0, nx = 0, nu = 0;
  int contactFrame[2] = {0, 3};
  double rect[4] = {0, 0, 0, 0};  // x_min, x_max, y_min, y_max
  std::vector<double> Qd, Rd, Qfd;
  double gPosZ = 0, gOri = 0, gLinVelZ = 0, gLinVelXY = 0, gAngVel = 0, gLinAccZ = 0, gLinAccXY = 0, gAngAcc = 0;
  double footW[18] = {0};
  double fricCoeff = 0, fricMu = 0, fricDelta = 0, fricReg = 25.0, fricShift = 1e-6;
  double momMu = 0, momDelta = 0;
  double jlMu = 0, jlDelta = 0;
  double collMu = 0, collDelta = 0, rFoot = 0, rKnee = 0;
  int armJoint[4] = {0, 0, 0, 0};
};

struct WbNode {  // per shooting node reference data (the per-instance inputs of b200sqp_upload_instances)
  uint8_t contact[2] = {1, 1};
  double swing[2][3] = {{0, 0, 0}, {0, 0, 0}};  // swing-z position, velocity, acceleration reference
  double impact[2] = {1, 1};                    // impact proximity factor
  double armPhase = 0;                          // sin(2 pi (phase - 0.15))
  Vec xref;                                     // TargetTrajectories::getDesiredState(t)
};

// ---- penalties ----------------------------------------------------------------------------------------------------------
struct Pen {
  double v, d1, d2;
};
inline Pen relaxedBarrier(double mu, double delta, double h) {
  if (h > delta) return {-mu * std::log(h), -mu / h, mu / (h * h)};
  const double dh = (h - 2.0 * delta) / delta;
  return {mu * (-std::log(delta) + 0.5 * dh * dh - 0.5), mu * ((h - 2.0 * delta) / (delta * delta)), mu / (delta * delta)};
}
inline Pen pwPolyBarrier(double mu, double delta, double h) {
  if (h <= 0) return {mu * (0.5 * h * h - delta * h / 2 + delta * delta / 6), mu * (h - delta / 2), mu};
  if (h < delta)
    return {mu * (-h * h * h / (6 * delta) + 0.5 * h * h - delta * h / 2 + delta * delta / 6), mu * (-h * h / (2 * delta) + h - delta / 2),
            mu * (-h / delta + 1)};
  return {0.0, 0.0, 0.0};
}

// ---- dynamics ---------------------------------------------------------------------------------------------------------------
// Base acceleration with the reference's block-diagonal base-inertia inverse.  M_bj*qdd_j + nle_b is evaluated as one
// RNEA with zero base acceleration (identical to crba/nonLinearEffects products; checked by wbBaseAccelerationLiteral).
template <class S>
void wbBaseAcceleration(const WbParams& P, const S* x, const S* u, S* qddb, KinData<S>* kinOut = nullptr) {
  const RobotModel& m = P.model;
  const int nv = m.nv();
  const S* q = x;
  const S* v = x + nv;
  std::vector<S> a(nv, S(0.0));
  for (int j = 0; j < m.nj; ++j) a[6 + j] = u[12 + j];
  std::vector<S> tau(nv);
  rnea(m, q, v, a.data(), tau.data());
  // contact wrenches through the base columns of the contact-frame Jacobians
  std::vector<S> zero(nv, S(0.0));
  KinData<S> kd;
  forwardKinematics(m, q, zero.data(), static_cast<const S*>(nullptr), V3<S>(), kd);
  S ext[6] = {S(0.0), S(0.0), S(0.0), S(0.0), S(0.0), S(0.0)};
  for (int c = 0; c < 2; ++c) {
    const FrameKin<S> fk = frameKinematics(m, kd, P.contactFrame[c]);
    S o[6];
    baseJacobianTransposeWrench(kd, fk.pos, V3<S>(u[6 * c], u[6 * c + 1], u[6 * c + 2]), V3<S>(u[6 * c + 3], u[6 * c + 4], u[6 * c + 5]), o);
    for (int k = 0; k < 6; ++k) ext[k] = ext[k] + o[k];
  }
  // M_bb: linear block = total mass * I (translation joint in world axes); angular block = S^T Ic S with Ic the composite
  // rotational inertia about the base origin in base axes.
  const int nb = m.nj + 1;
  std::vector<I6<S>> Y(nb);
  for (int i = 0; i < nb; ++i) Y[i] = bodyInertia6<S>(m.inertia[i]);
  for (int i = nb - 1; i >= 1; --i) {
    const I6<S> Yp = transformInertia(kd.liMi[i], Y[i]);
    for (int k = 0; k < 36; ++k) Y[m.parent[i]].m[k] = Y[m.parent[i]].m[k] + Yp.m[k];
  }
  // M_lin = R (m I) R^T = m I ; read it from the composite to stay literal
  M3<S> Mlin, Mang, Ic;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      Ic(i, j) = Y[0](3 + i, 3 + j);
      Mlin(i, j) = Y[0](i, j);
    }
  Mlin = kd.liMi[0].R * Mlin * transpose(kd.liMi[0].R);
  Mang = transpose(kd.Szyx) * Ic * kd.Szyx;
  S r[6];
  for (int k = 0; k < 6; ++k) r[k] = ext[k] - tau[k];
  auto solve3 = [](const M3<S>& A, const S* b, S* out) {
    // inverse by cofactors (Eigen's 3x3 inverse is the cofactor formula as well)
    const S c00 = A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1), c01 = A(1, 2) * A(2, 0) - A(1, 0) * A(2, 2), c02 = A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0);
    const S det = A(0, 0) * c00 + A(0, 1) * c01 + A(0, 2) * c02;
    const S i00 = c00 / det, i01 = (A(0, 2) * A(2, 1) - A(0, 1) * A(2, 2)) / det, i02 = (A(0, 1) * A(1, 2) - A(0, 2) * A(1, 1)) / det;
    const S i10 = c01 / det, i11 = (A(0, 0) * A(2, 2) - A(0, 2) * A(2, 0)) / det, i12 = (A(0, 2) * A(1, 0) - A(0, 0) * A(1, 2)) / det;
    const S i20 = c02 / det, i21 = (A(0, 1) * A(2, 0) - A(0, 0) * A(2, 1)) / det, i22 = (A(0, 0) * A(1, 1) - A(0, 1) * A(1, 0)) / det;
    out[0] = i00 * b[0] + i01 * b[1] + i02 * b[2];
    out[1] = i10 * b[0] + i11 * b[1] + i12 * b[2];
    out[2] = i20 * b[0] + i21 * b[1] + i22 * b[2];
  };
  solve3(Mlin, r, qddb);
  solve3(Mang, r + 3, qddb + 3);
  if (kinOut) *kinOut = kd;
}

// Literal restatement (crba + nonLinearEffects + Jacobian products) used to validate the fused version above.
inline void wbBaseAccelerationLiteral(const WbParams& P, const double* x, const double* u, double* qddb) {
  const RobotModel& m = P.model;
  const int nv = m.nv();
  std::vector<double> M;
  crba<double>(m, x, M);
  std::vector<double> nle(nv);
  rnea<double>(m, x, x + nv, static_cast<const double*>(nullptr), nle.data());
  std::vector<double> zero(nv, 0.0);
  KinData<double> kd;
  forwardKinematics<double>(m, x, zero.data(), nullptr, V3<double>(), kd);
  double ext[6] = {0, 0, 0, 0, 0, 0};
  for (int c = 0; c < 2; ++c) {
    const FrameKin<double> fk = frameKinematics(m, kd, P.contactFrame[c]);
    double o[6];
    baseJacobianTransposeWrench(kd, fk.pos, V3<double>(u[6 * c], u[6 * c + 1], u[6 * c + 2]), V3<double>(u[6 * c + 3], u[6 * c + 4], u[6 * c + 5]), o);
    for (int k = 0; k < 6; ++k) ext[k] += o[k];
  }
  double inter[6];
  for (int i = 0; i < 6; ++i) {
    double s = -nle[i] + ext[i];
    for (int j = 0; j < m.nj; ++j) s -= M[static_cast<size_t>(i) * nv + 6 + j] * u[12 + j];
    inter[i] = s;
  }
  Mat Ml(3, 3), Ma(3, 3);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      Ml(i, j) = M[static_cast<size_t>(i) * nv + j];
      Ma(i, j) = M[static_cast<size_t>(3 + i) * nv + 3 + j];
    }
  const Mat Mli = inverse(Ml), Mai = inverse(Ma);
  for (int i = 0; i < 3; ++i) {
    qddb[i] = Mli(i, 0) * inter[0] + Mli(i, 1) * inter[1] + Mli(i, 2) * inter[2];
    qddb[3 + i] = Mai(i, 0) * inter[3] + Mai(i, 1) * inter[4] + Mai(i, 2) * inter[5];
  }
}

template <class S>
void wbFlowMap(const WbParams& P, const S* x, const S* u, S* xdot) {
  const int nv = P.model.nv();
  S qddb[6];
  wbBaseAcceleration(P, x, u, qddb);
  for (int i = 0; i < nv; ++i) xdot[i] = x[nv + i];
  for (int i = 0; i < 6; ++i) xdot[nv + i] = qddb[i];
  for (int j = 0; j < P.model.nj; ++j) xdot[nv + 6 + j] = u[12 + j];
}

// computeJointTorques (humanoid_common_mpc/src/pinocchio_model/DynamicsHelperFunctions.cpp:232-270): tau_j = M_j [qdd_b; qdd_j] + nle_j - (J' W)_j
// with qdd_b from computeBaseAcceleration; M a + nle is one RNEA, the joint rows of J' W are axis . (M + (p_f - p_joint) x F) along each leg.
inline void wbJointTorques(const WbParams& P, const double* x, const double* u, double* tau, double* qddbOut = nullptr) {
  const RobotModel& m = P.model;
  const int nv = m.nv();
  double qddb[6];
  KinData<double> kd;
  wbBaseAcceleration<double>(P, x, u, qddb, &kd);   // kd: placements (zero-velocity pass)
  std::vector<double> a(nv, 0.0), full(nv);
  for (int k = 0; k < 6; ++k) a[k] = qddb[k];
  for (int j = 0; j < m.nj; ++j) a[6 + j] = u[12 + j];
  rnea<double>(m, x, x + nv, a.data(), full.data());
  for (int c = 0; c < 2; ++c) {
    const FrameKin<double> fk = frameKinematics(m, kd, P.contactFrame[c]);
    const V3<double> F(u[6 * c], u[6 * c + 1], u[6 * c + 2]), Mo(u[6 * c + 3], u[6 * c + 4], u[6 * c + 5]);
    for (int b = m.frameBody[P.contactFrame[c]]; b > 0; b = m.parent[b]) {
      const V3<double> ax = kd.oMi[b].R * V3<double>(m.axis[b][0], m.axis[b][1], m.axis[b][2]);
      full[5 + b] -= dot(ax, Mo + cross(fk.pos - kd.oMi[b].p, F));
    }
  }
  for (int j = 0; j < m.nj; ++j) tau[j] = full[6 + j];
  if (qddbOut) std::copy(qddb, qddb + 6, qddbOut);
}

// ---- end-effector quantities at (x,u) -------------------------------------------------------------------------------------------
template <class S>
struct FootState {
  FrameKin<S> fk;       // position, rotation, twist, classical acceleration (LOCAL_WORLD_ALIGNED)
  V3<S> oriErr;         // rotationMatrixDistanceToPlane(R, e_z)
};
// getQuaternionFromUnitVectors(R e_z, n) then quaternionDistance(q, Identity) = -q.vec  (RotationTransforms.h:98-113,51-53,396-405)
template <class S>
V3<S> orientationErrorToPlane(const M3<S>& R) {
  const V3<S> a(R(0, 2), R(1, 2), R(2, 2));
  const V3<S> b(S(0.0), S(0.0), S(1.0));
  const V3<S> c = cross(a, b);
  const S w = S(1.0) + dot(a, b);
  const S norm = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2] + w * w);
  return {-(c[0] / norm), -(c[1] / norm), -(c[2] / norm)};
}

template <class S>
struct WbEval {  // everything a node needs, evaluated once on scalar type S
  std::vector<S> xdot;        // flow map
  FootState<S> foot[2];
  std::vector<V3<S>> framePos;  // all operational frames
  S qddb[6];
};

template <class S>
void wbEvaluate(const WbParams& P, const S* x, const S* u, WbEval<S>& e) {
  const RobotModel& m = P.model;
  const int nv = m.nv();
  wbBaseAcceleration(P, x, u, e.qddb);
  e.xdot.assign(2 * nv, S(0.0));
  for (int i = 0; i < nv; ++i) e.xdot[i] = x[nv + i];
  for (int i = 0; i < 6; ++i) e.xdot[nv + i] = e.qddb[i];
  for (int j = 0; j < m.nj; ++j) e.xdot[nv + 6 + j] = u[12 + j];
  // forwardKinematics(q, v, a) with a = generalized accelerations, no gravity (PinocchioEndEffectorDynamicsCppAd.cpp:761-779)
  KinData<S> kd;
  forwardKinematics(m, x, x + nv, e.xdot.data() + nv, V3<S>(), kd);
  for (int c = 0; c < 2; ++c) {
    e.foot[c].fk = frameKinematics(m, kd, P.contactFrame[c]);
    e.foot[c].oriErr = orientationErrorToPlane(e.foot[c].fk.R);
  }
  e.framePos.resize(m.frameBody.size());
  for (size_t f = 0; f < m.frameBody.size(); ++f) e.framePos[f] = frameKinematics(m, kd, static_cast<int>(f)).pos;
}

// state-input equality constraints in the collection's order: per foot {zeroWrench | stance acceleration | swing-z}
template <class S>
void wbEqConstraints(const WbParams& P, const WbNode& nd, const S* u, const WbEval<S>& e, std::vector<S>& g) {
  g.clear();
  for (int c = 0; c < 2; ++c) {
    const FootState<S>& f = e.foot[c];
    if (!nd.contact[c]) {  // ZeroWrenchConstraint
      for (int k = 0; k < 6; ++k) g.push_back(u[6 * c + k]);
    }
    if (nd.contact[c]) {  // ZeroAccelerationConstraintCppAd: b + Ax [p; oriErr] + Av twist + Aa acc
      const double Av[6] = {P.gLinVelXY, P.gLinVelXY, P.gLinVelZ, P.gAngVel, P.gAngVel, P.gAngVel};
      const double Aa[6] = {P.gLinAccXY, P.gLinAccXY, P.gLinAccZ, P.gAngAcc, P.gAngAcc, P.gAngAcc};
      for (int k = 0; k < 3; ++k) {
        S val = S(Av[k]) * f.fk.vlin[k] + S(Aa[k]) * f.fk.alin[k];
        if (k == 2) val = val + S(P.gPosZ) * f.fk.pos[2];
        g.push_back(val);
      }
      for (int k = 0; k < 3; ++k) g.push_back(S(P.gOri) * f.oriErr[k] + S(Av[3 + k]) * f.fk.vang[k] + S(Aa[3 + k]) * f.fk.aang[k]);
    }
    if (!nd.contact[c]) {  // SwingLegVerticalConstraintCppAd (WBMpcPreComputation.cpp:92-103)
      const double b = -P.gLinVelZ * nd.swing[c][1] - P.gLinAccZ * nd.swing[c][2] - P.gPosZ * nd.swing[c][0];
      g.push_back(S(b) + S(P.gPosZ) * f.fk.pos[2] + S(P.gLinVelZ) * f.fk.vlin[2] + S(P.gLinAccZ) * f.fk.alin[2]);
    }
  }
}

// EndEffectorDynamicsFootCost residual (18) of a swing foot
template <class S>
void wbFootCostResidual(const WbParams& P, const WbNode& nd, int c, const WbEval<S>& e, S* res) {
  const FootState<S>& f = e.foot[c];
  const V3<S> zero;
  const V3<S>* parts[6] = {&zero, &f.oriErr, &f.fk.vlin, &f.fk.vang, &f.fk.alin, &f.fk.aang};
  for (int b = 0; b < 6; ++b)
    for (int k = 0; k < 3; ++k) res[3 * b + k] = (*parts[b])[k] * S(std::sqrt(P.footW[3 * b + k]) * nd.impact[c]);
}

// ContactMomentXY constraint values (4) of a stance foot
template <class S>
void wbMomentXY(const WbParams& P, int c, const S* u, const WbEval<S>& e, S* h) {
  const M3<S>& R = e.foot[c].fk.R;
  const V3<S> lf = tmul(R, V3<S>(u[6 * c], u[6 * c + 1], u[6 * c + 2]));
  const V3<S> lm = tmul(R, V3<S>(u[6 * c + 3], u[6 * c + 4], u[6 * c + 5]));
  h[0] = lm[0] - S(P.rect[2]) * lf[2];
  h[1] = -lm[0] + S(P.rect[3]) * lf[2];
  h[2] = -lm[1] - S(P.rect[0]) * lf[2];
  h[3] = lm[1] + S(P.rect[1]) * lf[2];
}

// FootCollisionConstraint values (16); frames: [0..2] left contact,p1,p2; [3..5] right; 6,7 ankles l,r; 8,9 knees l,r
template <class S>
void wbCollision(const WbParams& P, const WbEval<S>& e, S* h) {
  auto dist = [&](int a, int b) {
    const V3<S> d = e.framePos[a] - e.framePos[b];
    return sqrt(dot(d, d));
  };
  const S mf(2.0 * P.rFoot), mk(2.0 * P.rKnee);
  const int fl = 0, fl1 = 1, fl2 = 2, fr = 3, fr1 = 4, fr2 = 5, al = 6, ar = 7, kl = 8, kr = 9;
  h[0] = dist(fl1, fr1) - mf;
  h[1] = dist(fl1, fr2) - mf;
  h[2] = dist(fl2, fr1) - mf;
  h[3] = dist(fl2, fr2) - mf;
  h[4] = dist(fl, fr1) - mf;
  h[5] = dist(fl, fr2) - mf;
  h[6] = dist(fr, fl1) - mf;
  h[7] = dist(fr, fl2) - mf;
  h[8] = dist(fl, fr) - mf;
  h[9] = dist(kl, kr) - mk;
  h[10] = dist(fl, ar) - mf;
  h[11] = dist(fl1, ar) - mf;
  h[12] = dist(fl2, ar) - mf;
  h[13] = dist(fr, al) - mf;
  h[14] = dist(fr1, al) - mf;
  h[15] = dist(fr2, al) - mf;
}

// ---- the OCP -----------------------------------------------------------------------------------------------------------------------
class WbOcp : public Ocp {
 public:
  static constexpr int ND = 93;  // tangent directions: x (58) + u (35) for G1; model must satisfy nx+nu <= ND
  using D = Dual<ND>;
  WbParams P;
  std::vector<WbNode> nodes;

  explicit WbOcp(WbParams p) : P(std::move(p)) {
    nx = P.nx;
    nu = P.nu;
    if (nx + nu > ND) throw std::runtime_error("WbOcp: nx+nu exceeds the compiled tangent width");
  }

  // xNominal with the arm-swing reference evaluated at the current yaw (treated as a constant by the quadratic cost)
  Vec xNominal(const WbNode& nd, const Vec& x) const {
    Vec xn = nd.xref;
    const int nv = P.model.nv();
    const double yaw = x[3];
    const double localVx = std::cos(yaw) * xn[nv + 0] + std::sin(yaw) * xn[nv + 1];
    const double g = nd.armPhase * localVx;
    xn[6 + P.armJoint[0]] += -0.15 * g;
    xn[6 + P.armJoint[1]] += 0.15 * g;
    xn[6 + P.armJoint[2]] += -0.15 * g;
    xn[6 + P.armJoint[3]] += 0.15 * g;
    return xn;
  }
  Vec uNominal(const WbNode& nd) const {
    Vec un = vzero(nu);
    const int ns = nd.contact[0] + nd.contact[1];
    if (ns > 0) {
      const double fz = P.model.totalMass() * 9.81 / ns;
      for (int c = 0; c < 2; ++c)
        if (nd.contact[c]) un[6 * c + 2] = fz;
    }
    return un;
  }

  Vec flowMap(int, double, const Vec& x, const Vec& u) override {
    Vec xd(nx);
    wbFlowMap<double>(P, x.data(), u.data(), xd.data());
    return xd;
  }
  LinApprox flowMapLin(int, double, const Vec& x, const Vec& u) override {
    std::vector<D> xs, us;
    seed(x, u, xs, us);
    std::vector<D> xd(nx);
    wbFlowMap<D>(P, xs.data(), us.data(), xd.data());
    LinApprox l;
    l.f.resize(nx);
    l.dfdx = Mat(nx, nx);
    l.dfdu = Mat(nx, nu);
    for (int i = 0; i < nx; ++i) {
      l.f[i] = xd[i].v;
      for (int j = 0; j < nx; ++j) l.dfdx(i, j) = xd[i].d[j];
      for (int j = 0; j < nu; ++j) l.dfdu(i, j) = xd[i].d[nx + j];
    }
    return l;
  }

  double cost(int k, double, const Vec& x, const Vec& u) override {
    const WbNode& nd = nodes[k];
    WbEval<double> e;
    wbEvaluate<double>(P, x.data(), u.data(), e);
    double f = 0;
    const Vec dx = x - xNominal(nd, x), du = u - uNominal(nd);
    for (int i = 0; i < nx; ++i) f += 0.5 * P.Qd[i] * dx[i] * dx[i];
    for (int i = 0; i < nu; ++i) f += 0.5 * P.Rd[i] * du[i] * du[i];
    for (int c = 0; c < 2; ++c) {
      if (!nd.contact[c]) {
        double r[18];
        wbFootCostResidual<double>(P, nd, c, e, r);
        double s = 0;
        for (double v : r) s += v * v;
        f += 0.5 * s;
      } else {
        f += relaxedBarrier(P.fricMu, P.fricDelta, frictionCone(u.data() + 6 * c)).v;
        double h[4];
        wbMomentXY<double>(P, c, u.data(), e, h);
        for (double v : h) f += relaxedBarrier(P.momMu, P.momDelta, v).v;
      }
    }
    for (int j = 0; j < P.nj; ++j) {
      f += pwPolyBarrier(P.jlMu, P.jlDelta, P.model.qUpper[j] - x[6 + j]).v;
      f += pwPolyBarrier(P.jlMu, P.jlDelta, x[6 + j] - P.model.qLower[j]).v;
    }
    if (!(nd.contact[0] && nd.contact[1])) {
      double h[16];
      wbCollision<double>(P, e, h);
      for (double v : h) f += pwPolyBarrier(P.collMu, P.collDelta, v).v;
    }
    return f;
  }

  QuadApprox costQuad(int k, double, const Vec& x, const Vec& u) override {
    const WbNode& nd = nodes[k];
    std::vector<D> xs, us;
    seed(x, u, xs, us);
    WbEval<D> e;
    wbEvaluate<D>(P, xs.data(), us.data(), e);
    QuadApprox c;
    c.dfdxx = Mat(nx, nx);
    c.dfdux = Mat(nu, nx);
    c.dfduu = Mat(nu, nu);
    c.dfdx = vzero(nx);
    c.dfdu = vzero(nu);
    // (1) StateInputQuadraticCost
    const Vec dx = x - xNominal(nd, x), du = u - uNominal(nd);
    for (int i = 0; i < nx; ++i) {
      c.f += 0.5 * P.Qd[i] * dx[i] * dx[i];
      c.dfdx[i] += P.Qd[i] * dx[i];
      c.dfdxx(i, i) += P.Qd[i];
    }
    for (int i = 0; i < nu; ++i) {
      c.f += 0.5 * P.Rd[i] * du[i] * du[i];
      c.dfdu[i] += P.Rd[i] * du[i];
      c.dfduu(i, i) += P.Rd[i];
    }
    // helper: add penalty of a vector of dual-valued constraints (first-order constraint model)
    auto addPenalty = [&](const D* h, int n, const std::function<Pen(double)>& pen) {
      for (int r = 0; r < n; ++r) {
        const Pen p = pen(h[r].v);
        c.f += p.v;
        for (int i = 0; i < nx; ++i) c.dfdx[i] += p.d1 * h[r].d[i];
        for (int i = 0; i < nu; ++i) c.dfdu[i] += p.d1 * h[r].d[nx + i];
        if (p.d2 != 0.0) addOuter(c, h[r].d, p.d2);
      }
    };
    for (int cf = 0; cf < 2; ++cf) {
      if (!nd.contact[cf]) {  // (2) Gauss-Newton foot cost
        D r[18];
        wbFootCostResidual<D>(P, nd, cf, e, r);
        for (int rr = 0; rr < 18; ++rr) {
          c.f += 0.5 * r[rr].v * r[rr].v;
          for (int i = 0; i < nx; ++i) c.dfdx[i] += r[rr].v * r[rr].d[i];
          for (int i = 0; i < nu; ++i) c.dfdu[i] += r[rr].v * r[rr].d[nx + i];
          addOuter(c, r[rr].d, 1.0);
        }
      } else {
        // (3) friction cone, quadratic-order constraint (analytic, as in the reference)
        const double* F = u.data() + 6 * cf;
        const double Ft2 = F[0] * F[0] + F[1] * F[1] + P.fricReg, Ft = std::sqrt(Ft2), Ft32 = Ft * Ft2;
        const double h = frictionCone(F);
        const double dh[3] = {-F[0] / Ft, -F[1] / Ft, P.fricCoeff};
        double ddh[3][3] = {{-(F[1] * F[1] + P.fricReg) / Ft32, F[0] * F[1] / Ft32, 0}, {F[0] * F[1] / Ft32, -(F[0] * F[0] + P.fricReg) / Ft32, 0}, {0, 0, 0}};
        const Pen p = relaxedBarrier(P.fricMu, P.fricDelta, h);
        c.f += p.v;
        for (int i = 0; i < 3; ++i) c.dfdu[6 * cf + i] += p.d1 * dh[i];
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 3; ++j) c.dfduu(6 * cf + i, 6 * cf + j) += p.d2 * dh[i] * dh[j] + p.d1 * ddh[i][j];
        for (int i = 0; i < nu; ++i) c.dfduu(i, i) += p.d1 * (-P.fricShift);
        for (int i = 0; i < nx; ++i) c.dfdxx(i, i) += p.d1 * (-P.fricShift);
        // (4) contact moment XY
        D hm[4];
        wbMomentXY<D>(P, cf, us.data(), e, hm);
        addPenalty(hm, 4, [&](double v) { return relaxedBarrier(P.momMu, P.momDelta, v); });
      }
    }
    // (5) joint limits
    for (int j = 0; j < P.nj; ++j) {
      const Pen pu = pwPolyBarrier(P.jlMu, P.jlDelta, P.model.qUpper[j] - x[6 + j]);
      const Pen pl = pwPolyBarrier(P.jlMu, P.jlDelta, x[6 + j] - P.model.qLower[j]);
      c.f += pu.v + pl.v;
      c.dfdx[6 + j] += pl.d1 - pu.d1;
      c.dfdxx(6 + j, 6 + j) += pl.d2 + pu.d2;
    }
    // (6) foot collision
    if (!(nd.contact[0] && nd.contact[1])) {
      D h[16];
      wbCollision<D>(P, e, h);
      addPenalty(h, 16, [&](double v) { return pwPolyBarrier(P.collMu, P.collDelta, v); });
    }
    return c;
  }

  double finalCost(int k, double, const Vec& x) override {
    const Vec dx = x - nodes[k].xref;
    double f = 0;
    for (int i = 0; i < nx; ++i) f += 0.5 * P.Qfd[i] * dx[i] * dx[i];
    return f;
  }
  QuadApprox finalCostQuad(int k, double t, const Vec& x) override {
    QuadApprox c;
    c.f = finalCost(k, t, x);
    c.dfdxx = Mat(nx, nx);
    c.dfdx = vzero(nx);
    for (int i = 0; i < nx; ++i) {
      c.dfdxx(i, i) = P.Qfd[i];
      c.dfdx[i] = P.Qfd[i] * (x[i] - nodes[k].xref[i]);
    }
    return c;
  }

  Vec eqConstraint(int k, double, const Vec& x, const Vec& u) override {
    WbEval<double> e;
    wbEvaluate<double>(P, x.data(), u.data(), e);
    std::vector<double> g;
    wbEqConstraints<double>(P, nodes[k], u.data(), e, g);
    return g;
  }
  LinApprox eqConstraintLin(int k, double, const Vec& x, const Vec& u) override {
    std::vector<D> xs, us;
    seed(x, u, xs, us);
    WbEval<D> e;
    wbEvaluate<D>(P, xs.data(), us.data(), e);
    std::vector<D> g;
    wbEqConstraints<D>(P, nodes[k], us.data(), e, g);
    LinApprox l;
    const int nc = static_cast<int>(g.size());
    l.f.resize(nc);
    l.dfdx = Mat(nc, nx);
    l.dfdu = Mat(nc, nu);
    for (int r = 0; r < nc; ++r) {
      l.f[r] = g[r].v;
      for (int j = 0; j < nx; ++j) l.dfdx(r, j) = g[r].d[j];
      for (int j = 0; j < nu; ++j) l.dfdu(r, j) = g[r].d[nx + j];
    }
    return l;
  }

  double frictionCone(const double* F) const {
    return P.fricCoeff * F[2] - std::sqrt(F[0] * F[0] + F[1] * F[1] + P.fricReg);
  }

 private:
  void seed(const Vec& x, const Vec& u, std::vector<D>& xs, std::vector<D>& us) const {
    xs.resize(nx);
    us.resize(nu);
    for (int i = 0; i < nx; ++i) xs[i] = D::variable(x[i], i);
    for (int i = 0; i < nu; ++i) us[i] = D::variable(u[i], nx + i);
  }
  void addOuter(QuadApprox& c, const double* d, double w) const {
    for (int j = 0; j < nx; ++j) {
      const double wj = w * d[j];
      if (wj == 0.0) continue;
      for (int i = 0; i < nx; ++i) c.dfdxx(i, j) += d[i] * wj;
      for (int i = 0; i < nu; ++i) c.dfdux(i, j) += d[nx + i] * wj;
    }
    for (int j = 0; j < nu; ++j) {
      const double wj = w * d[nx + j];
      if (wj == 0.0) continue;
      for (int i = 0; i < nu; ++i) c.dfduu(i, j) += d[nx + i] * wj;
    }
  }
};

}  // namespace orc
