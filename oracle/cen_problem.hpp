// ORACLE (test infrastructure only; never linked into libb200sqp.so) -- CPU restatement of the Unitree G1 centroidal OCP terms.
//
// x = [h/m (lin 3, ang 3); base position 3; Euler ZYX 3; joints nj], u = [W_left(6); W_right(6); joint velocities nj]
//   (humanoid_nmpc/humanoid_centroidal_mpc/include/humanoid_centroidal_mpc/common/CentroidalMpcRobotModel.h:52-160)
// Terms and their reference sources (paths relative to /root/reference/humanoid_nmpc/), in the order of
// CentroidalMpcInterface::setupOptimalControlProblem (humanoid_centroidal_mpc/src/CentroidalMpcInterface.cpp:150-237):
//   dynamics              cen_dynamics.hpp (PinocchioCentroidalDynamicsAD)
//   quadratic tracking    humanoid_common_mpc/src/cost/StateInputQuadraticCost.cpp:67-78, reference_manager/SwitchedModelReferenceManager.cpp:110-135
//   terminal cost         humanoid_common_mpc/src/HumanoidCostConstraintFactory.cpp:218-228
//   task-space link cost  humanoid_common_mpc/src/cost/EndEffectorKinematicsQuadraticCost.cpp:76-128, EndEffectorKinematicCostHelpers.cpp:113-122
//                         (quaternionDistance, matrixToQuaternion: ocs2_robotic_tools/common/RotationTransforms.h:51-53,215-245)
//   ICP cost              humanoid_centroidal_mpc/src/cost/ICPCost.cpp:78-108 (weight 0 in the shipped task.info)
//   joint limits          humanoid_common_mpc/src/constraint/JointLimitsSoftConstraint.cpp:69-100        (stateSoftConstraint)
//   foot collision        humanoid_common_mpc/src/constraint/FootCollisionConstraint.cpp:80-144           (stateSoftConstraint)
//   friction cone         humanoid_common_mpc/src/constraint/FrictionForceConeConstraint.cpp:80-224
//   contact moment XY     humanoid_common_mpc/src/constraint/ContactMomentXYConstraintCppAd.cpp:86-104
//   zero wrench           humanoid_common_mpc/src/constraint/ZeroWrenchConstraint.cpp:59-84
//   zero velocity         humanoid_centroidal_mpc/src/constraint/ZeroVelocityConstraintCppAd.cpp, humanoid_common_mpc/src/constraint/
//                         EndEffectorKinematicsTwistConstraint.cpp:79-131, config CentroidalMpcInterface.cpp:246-268
//   normal velocity       humanoid_centroidal_mpc/src/constraint/NormalVelocityConstraintCppAd.cpp:62-85, humanoid_common_mpc/src/constraint/
//                         EndEffectorKinematicsLinearVelConstraint.cpp:71-111, config humanoid_common_mpc/src/HumanoidPreComputation.cpp:100-121
//   foot tracking cost    humanoid_centroidal_mpc/src/cost/CentroidalMpcEndEffectorFootCost.cpp:91-147
//   external torque cost  humanoid_common_mpc/src/cost/ExternalTorqueQuadraticCostAD.cpp:84-131
//   initializer           humanoid_centroidal_mpc/src/initialization/CentroidalWeightCompInitializer.cpp:66-74 (host side, references.py)
#pragma once
#include "cen_dynamics.hpp"
#include "wb_problem.hpp"

namespace orc {

struct CenParams {
  WbParams base;  // tree, frames, contact rectangle, Q/R/Qf (nx = nu = 12 + nj), gPosZ, gOri, footW[0..11], barrier parameters, arm joints
  int torsoFrame = 10;      // index into model.frameBody / frameP of the task-space link
  double torsoR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};  // rotation of the link frame in its body's joint frame (row-major)
  double torsoW[12] = {0};
  double icpW = 0;
  int tqJoint[2][6] = {{0}};
  double tqW[2][6] = {{0}};
};

// matrixToQuaternion, CppAD flavour (RotationTransforms.h:215-245): the branches are CondExp operations there, i.e. evaluated on values.
template <class S>
void matrixToQuaternion(const M3<S>& R, S q[4] /* x y z w */) {
  const double r00 = value(R(0, 0)), r11 = value(R(1, 1)), r22 = value(R(2, 2));
  const bool gt = r00 > r11, lt = r00 < -r11, neg = r22 < 0.0;
  const S one(1.0);
  const S t1 = gt ? one + R(0, 0) - R(1, 1) - R(2, 2) : one - R(0, 0) + R(1, 1) - R(2, 2);
  const S t2 = lt ? one - R(0, 0) - R(1, 1) + R(2, 2) : one + R(0, 0) + R(1, 1) + R(2, 2);
  const S t = neg ? t1 : t2;
  const S x1 = gt ? t : R(1, 0) + R(0, 1), x2 = lt ? R(0, 2) + R(2, 0) : R(2, 1) - R(1, 2);
  const S y1 = gt ? R(1, 0) + R(0, 1) : t, y2 = lt ? R(2, 1) + R(1, 2) : R(0, 2) - R(2, 0);
  const S z1 = gt ? R(0, 2) + R(2, 0) : R(2, 1) + R(1, 2), z2 = lt ? t : R(1, 0) - R(0, 1);
  const S w1 = gt ? R(2, 1) - R(1, 2) : R(0, 2) - R(2, 0), w2 = lt ? R(1, 0) - R(0, 1) : t;
  const S sc = S(0.5) / sqrt(t);
  q[0] = (neg ? x1 : x2) * sc;
  q[1] = (neg ? y1 : y2) * sc;
  q[2] = (neg ? z1 : z2) * sc;
  q[3] = (neg ? w1 : w2) * sc;
}
// quaternionDistance(q, qRef) = q.w qRef.vec - qRef.w q.vec + q.vec x qRef.vec   (RotationTransforms.h:51-53)
template <class S>
V3<S> quaternionDistance(const S q[4], const double qr[4]) {
  const V3<S> qv(q[0], q[1], q[2]);
  const S r0 = S(qr[0]), r1 = S(qr[1]), r2 = S(qr[2]);
  const V3<S> rv(r0, r1, r2);
  const V3<S> c = cross(qv, rv);
  return {q[3] * rv[0] - S(qr[3]) * qv[0] + c[0], q[3] * rv[1] - S(qr[3]) * qv[1] + c[1], q[3] * rv[2] - S(qr[3]) * qv[2] + c[2]};
}

template <class S>
struct CenEval {
  std::vector<S> xdot;  // flow map; xdot[6 .. 6+nv) = generalized velocities (getPinocchioJointVelocity)
  FrameKin<S> foot[2];
  V3<S> oriErr[2];  // rotationMatrixDistanceToPlane(R, e_z)
  FrameKin<S> torso;
  S torsoQuat[4];
  std::vector<V3<S>> framePos;
  V3<S> com;
  KinData<S> kin;
};

template <class S>
void cenEvaluate(const CenParams& P, const S* x, const S* u, CenEval<S>& e) {
  const RobotModel& m = P.base.model;
  const int nx = 12 + m.nj;
  e.xdot.assign(nx, S(0.0));
  centroidalFlowMap<S>(m, P.base.contactFrame, x, u, e.xdot.data());
  // forwardKinematics(q, v) with v = getGeneralizedVelocities(state, input)
  forwardKinematics<S>(m, x + 6, e.xdot.data() + 6, static_cast<const S*>(nullptr), V3<S>(), e.kin);
  for (int c = 0; c < 2; ++c) {
    e.foot[c] = frameKinematics(m, e.kin, P.base.contactFrame[c]);
    e.oriErr[c] = orientationErrorToPlane(e.foot[c].R);
  }
  e.torso = frameKinematics(m, e.kin, P.torsoFrame);
  M3<S> Rf;
  for (int k = 0; k < 9; ++k) Rf.m[k] = S(P.torsoR[k]);
  e.torso.R = e.torso.R * Rf;
  matrixToQuaternion(e.torso.R, e.torsoQuat);
  e.framePos.resize(m.frameBody.size());
  for (size_t f = 0; f < m.frameBody.size(); ++f) e.framePos[f] = frameKinematics(m, e.kin, static_cast<int>(f)).pos;
  e.com = centerOfMass(m, e.kin);
}

// state-input equality constraints in the collection's order: per foot {zeroWrench | zeroVelocity | normalVelocity}
template <class S>
void cenEqConstraints(const CenParams& P, const WbNode& nd, const S* u, const CenEval<S>& e, std::vector<S>& g) {
  g.clear();
  const WbParams& B = P.base;
  for (int c = 0; c < 2; ++c) {
    const FrameKin<S>& f = e.foot[c];
    if (!nd.contact[c])
      for (int k = 0; k < 6; ++k) g.push_back(u[6 * c + k]);
    if (nd.contact[c]) {  // Ax [p; oriErr] + Av twist with Ax(2,2) = positionErrorGain_z, Ax(3:6,3:6) = orientationErrorGain I, Av = I
      g.push_back(f.vlin[0]);
      g.push_back(f.vlin[1]);
      g.push_back(f.vlin[2] + S(B.gPosZ) * f.pos[2]);
      for (int k = 0; k < 3; ++k) g.push_back(f.vang[k] + S(B.gOri) * e.oriErr[c][k]);
    }
    if (!nd.contact[c]) {  // b = -zVelRef - gPosZ zPosRef, Av = e_z', Ax = gPosZ e_z'
      const double b = -nd.swing[c][1] - B.gPosZ * nd.swing[c][0];
      g.push_back(S(b) + f.vlin[2] + S(B.gPosZ) * f.pos[2]);
    }
  }
}

// CentroidalMpcEndEffectorFootCost residual (12): reference position 0, plane normal e_z, zero reference velocities
template <class S>
void cenFootResidual(const CenParams& P, const WbNode& nd, int c, const CenEval<S>& e, S* r) {
  const FrameKin<S>& f = e.foot[c];
  for (int k = 0; k < 3; ++k) {
    r[k] = f.pos[k] * S(std::sqrt(P.base.footW[k]));
    r[3 + k] = e.oriErr[c][k] * S(std::sqrt(P.base.footW[3 + k]));
    r[6 + k] = f.vlin[k] * S(nd.impact[c] * std::sqrt(P.base.footW[6 + k]));
    r[9 + k] = f.vang[k] * S(std::sqrt(P.base.footW[9 + k]));
  }
}

// reference cost element of the task-space link at (xRef, uRef = 0): position 3, quaternion xyzw 4, linear velocity 3, angular velocity 3
inline void cenTorsoReference(const CenParams& P, const Vec& xref, double ref[13]) {
  Vec u0 = vzero(static_cast<int>(xref.size()));
  CenEval<double> e;
  cenEvaluate<double>(P, xref.data(), u0.data(), e);
  for (int k = 0; k < 3; ++k) {
    ref[k] = e.torso.pos[k];
    ref[7 + k] = e.torso.vlin[k];
    ref[10 + k] = e.torso.vang[k];
  }
  for (int k = 0; k < 4; ++k) ref[3 + k] = e.torsoQuat[k];
}
template <class S>
void cenTorsoResidual(const CenParams& P, const double ref[13], const CenEval<S>& e, S* r) {
  const V3<S> oe = quaternionDistance(e.torsoQuat, ref + 3);
  for (int k = 0; k < 3; ++k) {
    r[k] = (e.torso.pos[k] - S(ref[k])) * S(std::sqrt(P.torsoW[k]));
    r[3 + k] = oe[k] * S(std::sqrt(P.torsoW[3 + k]));
    r[6 + k] = (e.torso.vlin[k] - S(ref[7 + k])) * S(std::sqrt(P.torsoW[6 + k]));
    r[9 + k] = (e.torso.vang[k] - S(ref[10 + k])) * S(std::sqrt(P.torsoW[9 + k]));
  }
}

// ExternalTorqueQuadraticCostAD residual (6) of stance foot c: joint rows of J_ee' W (LOCAL_WORLD_ALIGNED), scaled by the mid-swing factor
// of the OTHER foot
template <class S>
void cenTorqueResidual(const CenParams& P, const WbNode& nd, int c, const S* u, const CenEval<S>& e, S* r) {
  const RobotModel& m = P.base.model;
  const V3<S> F(u[6 * c], u[6 * c + 1], u[6 * c + 2]), Mo(u[6 * c + 3], u[6 * c + 4], u[6 * c + 5]);
  const int fb = m.frameBody[P.base.contactFrame[c]];
  const double mid = 1.0 - nd.impact[1 - c];
  for (int i = 0; i < 6; ++i) {
    const int body = P.tqJoint[c][i] + 1;
    bool onPath = false;
    for (int b = fb; b > 0; b = m.parent[b]) onPath = onPath || (b == body);
    S tau(0.0);
    if (onPath) {
      const V3<S> a = e.kin.oMi[body].R * V3<S>(S(m.axis[body][0]), S(m.axis[body][1]), S(m.axis[body][2]));
      tau = dot(a, Mo + cross(e.foot[c].pos - e.kin.oMi[body].p, F));
    }
    r[i] = tau * S(std::sqrt(P.tqW[c][i]) * mid);
  }
}

template <class S>
void cenMomentXY(const CenParams& P, int c, const S* u, const CenEval<S>& e, S* h) {
  const M3<S>& R = e.foot[c].R;
  const V3<S> lf = tmul(R, V3<S>(u[6 * c], u[6 * c + 1], u[6 * c + 2]));
  const V3<S> lm = tmul(R, V3<S>(u[6 * c + 3], u[6 * c + 4], u[6 * c + 5]));
  h[0] = lm[0] - S(P.base.rect[2]) * lf[2];
  h[1] = -lm[0] + S(P.base.rect[3]) * lf[2];
  h[2] = -lm[1] - S(P.base.rect[0]) * lf[2];
  h[3] = lm[1] + S(P.base.rect[1]) * lf[2];
}

template <class S>
void cenCollision(const CenParams& P, const CenEval<S>& e, S* h) {
  WbEval<S> w;
  w.framePos = e.framePos;
  wbCollision<S>(P.base, w, h);
}

class CenOcp : public Ocp {
 public:
  static constexpr int ND = 70;
  using D = Dual<ND>;
  CenParams P;
  std::vector<WbNode> nodes;

  explicit CenOcp(CenParams p) : P(std::move(p)) {
    nx = nu = 12 + P.base.model.nj;
    if (nx + nu > ND) throw std::runtime_error("CenOcp: nx+nu exceeds the compiled tangent width");
  }

  Vec xNominal(const WbNode& nd, const Vec& x) const {
    Vec xn = nd.xref;
    const double yaw = x[9];
    const double localVx = std::cos(yaw) * xn[0] + std::sin(yaw) * xn[1];
    const double g = nd.armPhase * localVx;
    xn[12 + P.base.armJoint[0]] += -0.15 * g;
    xn[12 + P.base.armJoint[1]] += 0.15 * g;
    xn[12 + P.base.armJoint[2]] += -0.15 * g;
    xn[12 + P.base.armJoint[3]] += 0.15 * g;
    return xn;
  }
  Vec uNominal(const WbNode& nd) const {
    Vec un = vzero(nu);
    const int ns = nd.contact[0] + nd.contact[1];
    if (ns > 0) {
      const double fz = P.base.model.totalMass() * 9.81 / ns;
      for (int c = 0; c < 2; ++c)
        if (nd.contact[c]) un[6 * c + 2] = fz;
    }
    return un;
  }

  Vec flowMap(int, double, const Vec& x, const Vec& u) override {
    Vec xd(nx);
    centroidalFlowMap<double>(P.base.model, P.base.contactFrame, x.data(), u.data(), xd.data());
    return xd;
  }
  LinApprox flowMapLin(int, double, const Vec& x, const Vec& u) override {
    std::vector<D> xs, us;
    seed(x, u, xs, us);
    std::vector<D> xd(nx);
    centroidalFlowMap<D>(P.base.model, P.base.contactFrame, xs.data(), us.data(), xd.data());
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

  // Gauss-Newton residual stack of node k: torso (12), ICP (2), per foot tracking (12) and external torque (6, stance only)
  template <class S>
  void residuals(const WbNode& nd, const S* u, const CenEval<S>& e, std::vector<S>& r) const {
    r.clear();
    S buf[12];
    double ref[13];
    cenTorsoReference(P, nd.xref, ref);
    cenTorsoResidual<S>(P, ref, e, buf);
    r.insert(r.end(), buf, buf + 12);
    const double sw = std::sqrt(P.icpW);
    for (int k = 0; k < 2; ++k) r.push_back(((e.foot[0].pos[k] + e.foot[1].pos[k]) / S(2.0) - e.com[k]) * S(sw));
    for (int c = 0; c < 2; ++c) {
      cenFootResidual<S>(P, nd, c, e, buf);
      r.insert(r.end(), buf, buf + 12);
      if (nd.contact[c]) {
        cenTorqueResidual<S>(P, nd, c, u, e, buf);
        r.insert(r.end(), buf, buf + 6);
      }
    }
  }

  double cost(int k, double, const Vec& x, const Vec& u) override {
    const WbNode& nd = nodes[k];
    const WbParams& B = P.base;
    CenEval<double> e;
    cenEvaluate<double>(P, x.data(), u.data(), e);
    double f = 0;
    const Vec dx = x - xNominal(nd, x), du = u - uNominal(nd);
    for (int i = 0; i < nx; ++i) f += 0.5 * B.Qd[i] * dx[i] * dx[i];
    for (int i = 0; i < nu; ++i) f += 0.5 * B.Rd[i] * du[i] * du[i];
    std::vector<double> r;
    residuals<double>(nd, u.data(), e, r);
    for (double v : r) f += 0.5 * v * v;
    for (int c = 0; c < 2; ++c)
      if (nd.contact[c]) {
        f += relaxedBarrier(B.fricMu, B.fricDelta, frictionCone(u.data() + 6 * c)).v;
        double h[4];
        cenMomentXY<double>(P, c, u.data(), e, h);
        for (double v : h) f += relaxedBarrier(B.momMu, B.momDelta, v).v;
      }
    for (int j = 0; j < B.nj; ++j) {
      f += pwPolyBarrier(B.jlMu, B.jlDelta, B.model.qUpper[j] - x[12 + j]).v;
      f += pwPolyBarrier(B.jlMu, B.jlDelta, x[12 + j] - B.model.qLower[j]).v;
    }
    if (!(nd.contact[0] && nd.contact[1])) {
      double h[16];
      cenCollision<double>(P, e, h);
      for (double v : h) f += pwPolyBarrier(B.collMu, B.collDelta, v).v;
    }
    return f;
  }

  QuadApprox costQuad(int k, double, const Vec& x, const Vec& u) override {
    const WbNode& nd = nodes[k];
    const WbParams& B = P.base;
    std::vector<D> xs, us;
    seed(x, u, xs, us);
    CenEval<D> e;
    cenEvaluate<D>(P, xs.data(), us.data(), e);
    QuadApprox c;
    c.dfdxx = Mat(nx, nx);
    c.dfdux = Mat(nu, nx);
    c.dfduu = Mat(nu, nu);
    c.dfdx = vzero(nx);
    c.dfdu = vzero(nu);
    const Vec dx = x - xNominal(nd, x), du = u - uNominal(nd);
    for (int i = 0; i < nx; ++i) {
      c.f += 0.5 * B.Qd[i] * dx[i] * dx[i];
      c.dfdx[i] += B.Qd[i] * dx[i];
      c.dfdxx(i, i) += B.Qd[i];
    }
    for (int i = 0; i < nu; ++i) {
      c.f += 0.5 * B.Rd[i] * du[i] * du[i];
      c.dfdu[i] += B.Rd[i] * du[i];
      c.dfduu(i, i) += B.Rd[i];
    }
    std::vector<D> r;
    residuals<D>(nd, us.data(), e, r);
    for (const D& rr : r) {  // StateInputCostGaussNewtonAd (ocs2_core/src/cost/StateInputGaussNewtonCostAd.cpp:76-95)
      c.f += 0.5 * rr.v * rr.v;
      for (int i = 0; i < nx; ++i) c.dfdx[i] += rr.v * rr.d[i];
      for (int i = 0; i < nu; ++i) c.dfdu[i] += rr.v * rr.d[nx + i];
      addOuter(c, rr.d, 1.0);
    }
    auto addPenalty = [&](const D* h, int n, const std::function<Pen(double)>& pen) {
      for (int rI = 0; rI < n; ++rI) {
        const Pen p = pen(h[rI].v);
        c.f += p.v;
        for (int i = 0; i < nx; ++i) c.dfdx[i] += p.d1 * h[rI].d[i];
        for (int i = 0; i < nu; ++i) c.dfdu[i] += p.d1 * h[rI].d[nx + i];
        if (p.d2 != 0.0) addOuter(c, h[rI].d, p.d2);
      }
    };
    for (int cf = 0; cf < 2; ++cf)
      if (nd.contact[cf]) {
        const double* F = u.data() + 6 * cf;
        const double Ft2 = F[0] * F[0] + F[1] * F[1] + B.fricReg, Ft = std::sqrt(Ft2), Ft32 = Ft * Ft2;
        const double h = frictionCone(F);
        const double dh[3] = {-F[0] / Ft, -F[1] / Ft, B.fricCoeff};
        const double ddh[3][3] = {{-(F[1] * F[1] + B.fricReg) / Ft32, F[0] * F[1] / Ft32, 0}, {F[0] * F[1] / Ft32, -(F[0] * F[0] + B.fricReg) / Ft32, 0}, {0, 0, 0}};
        const Pen p = relaxedBarrier(B.fricMu, B.fricDelta, h);
        c.f += p.v;
        for (int i = 0; i < 3; ++i) c.dfdu[6 * cf + i] += p.d1 * dh[i];
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 3; ++j) c.dfduu(6 * cf + i, 6 * cf + j) += p.d2 * dh[i] * dh[j] + p.d1 * ddh[i][j];
        for (int i = 0; i < nu; ++i) c.dfduu(i, i) += p.d1 * (-B.fricShift);
        for (int i = 0; i < nx; ++i) c.dfdxx(i, i) += p.d1 * (-B.fricShift);
        D hm[4];
        cenMomentXY<D>(P, cf, us.data(), e, hm);
        addPenalty(hm, 4, [&](double v) { return relaxedBarrier(B.momMu, B.momDelta, v); });
      }
    for (int j = 0; j < B.nj; ++j) {
      const Pen pu = pwPolyBarrier(B.jlMu, B.jlDelta, B.model.qUpper[j] - x[12 + j]);
      const Pen pl = pwPolyBarrier(B.jlMu, B.jlDelta, x[12 + j] - B.model.qLower[j]);
      c.f += pu.v + pl.v;
      c.dfdx[12 + j] += pl.d1 - pu.d1;
      c.dfdxx(12 + j, 12 + j) += pl.d2 + pu.d2;
    }
    if (!(nd.contact[0] && nd.contact[1])) {
      D h[16];
      cenCollision<D>(P, e, h);
      addPenalty(h, 16, [&](double v) { return pwPolyBarrier(B.collMu, B.collDelta, v); });
    }
    return c;
  }

  double finalCost(int k, double, const Vec& x) override {
    const Vec dx = x - nodes[k].xref;
    double f = 0;
    for (int i = 0; i < nx; ++i) f += 0.5 * P.base.Qfd[i] * dx[i] * dx[i];
    return f;
  }
  QuadApprox finalCostQuad(int k, double t, const Vec& x) override {
    QuadApprox c;
    c.f = finalCost(k, t, x);
    c.dfdxx = Mat(nx, nx);
    c.dfdx = vzero(nx);
    for (int i = 0; i < nx; ++i) {
      c.dfdxx(i, i) = P.base.Qfd[i];
      c.dfdx[i] = P.base.Qfd[i] * (x[i] - nodes[k].xref[i]);
    }
    return c;
  }

  Vec eqConstraint(int k, double, const Vec& x, const Vec& u) override {
    CenEval<double> e;
    cenEvaluate<double>(P, x.data(), u.data(), e);
    std::vector<double> g;
    cenEqConstraints<double>(P, nodes[k], u.data(), e, g);
    return g;
  }
  LinApprox eqConstraintLin(int k, double, const Vec& x, const Vec& u) override {
    std::vector<D> xs, us;
    seed(x, u, xs, us);
    CenEval<D> e;
    cenEvaluate<D>(P, xs.data(), us.data(), e);
    std::vector<D> g;
    cenEqConstraints<D>(P, nodes[k], us.data(), e, g);
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
    return P.base.fricCoeff * F[2] - std::sqrt(F[0] * F[0] + F[1] * F[1] + P.base.fricReg);
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
