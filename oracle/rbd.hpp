// ORACLE (test infrastructure only) -- rigid-body algorithms templated on the scalar type.
//
// Pinocchio is an un-vendored apt dependency of the reference (ros-${ROS_DISTRO}-pinocchio, version unpinned).  The
// algorithms the hot path calls are restated here from their published definitions (Featherstone, "Rigid Body Dynamics
// Algorithms", 2008; Carpentier et al., "The Pinocchio C++ library", 2019), in Pinocchio's conventions:
//   spatial vectors are [linear; angular]; body-local ("LOCAL") recursion; SE3 (R,p) maps child to parent coordinates.
// Call sites in the reference:
//   pinocchio::crba / nonLinearEffects / computeFrameJacobian(LOCAL_WORLD_ALIGNED)
//       humanoid_nmpc/humanoid_wb_mpc/src/dynamics/DynamicsHelperFunctions.cpp:63-73
//   forwardKinematics(q,v,a), getFrameVelocity, getFrameClassicalAcceleration(LOCAL_WORLD_ALIGNED)
//       humanoid_nmpc/humanoid_wb_mpc/src/end_effector/PinocchioEndEffectorDynamicsCppAd.cpp:246-247,300-305,587-592,650-655,769-774
//   base joint = JointModelComposite(Translation, SphericalZYX)
//       humanoid_nmpc/humanoid_common_mpc/src/pinocchio_model/createPinocchioModel.cpp:60-67
// G1-level parity with Pinocchio is UNPINNED by the reference's tests; it is pinned here by
// finite differences and physical identities (tests/test_oracle_wb.py).
#pragma once
#include <array>
#include <cmath>
#include <vector>

#include "dual.hpp"

namespace orc {

template <class S>
struct V3 {
  S x[3];
  V3() : x{S(0.0), S(0.0), S(0.0)} {}
  V3(S a, S b, S c) : x{a, b, c} {}
  S& operator[](int i) { return x[i]; }
  const S& operator[](int i) const { return x[i]; }
};
template <class S>
V3<S> operator+(const V3<S>& a, const V3<S>& b) { return {a[0] + b[0], a[1] + b[1], a[2] + b[2]}; }
template <class S>
V3<S> operator-(const V3<S>& a, const V3<S>& b) { return {a[0] - b[0], a[1] - b[1], a[2] - b[2]}; }
template <class S>
V3<S> operator-(const V3<S>& a) { return {-a[0], -a[1], -a[2]}; }
template <class S>
V3<S> operator*(const S& s, const V3<S>& a) { return {s * a[0], s * a[1], s * a[2]}; }
template <class S>
V3<S> cross(const V3<S>& a, const V3<S>& b) { return {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]}; }
template <class S>
S dot(const V3<S>& a, const V3<S>& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

template <class S>
struct M3 {  // row-major
  S m[9];
  M3() { for (auto& e : m) e = S(0.0); }
  S& operator()(int i, int j) { return m[3 * i + j]; }
  const S& operator()(int i, int j) const { return m[3 * i + j]; }
  static M3 identity() {
    M3 r;
    r(0, 0) = r(1, 1) = r(2, 2) = S(1.0);
    return r;
  }
};
template <class S>
V3<S> operator*(const M3<S>& A, const V3<S>& v) {
  return {A(0, 0) * v[0] + A(0, 1) * v[1] + A(0, 2) * v[2], A(1, 0) * v[0] + A(1, 1) * v[1] + A(1, 2) * v[2],
          A(2, 0) * v[0] + A(2, 1) * v[1] + A(2, 2) * v[2]};
}
template <class S>
V3<S> tmul(const M3<S>& A, const V3<S>& v) {  // A^T v
  return {A(0, 0) * v[0] + A(1, 0) * v[1] + A(2, 0) * v[2], A(0, 1) * v[0] + A(1, 1) * v[1] + A(2, 1) * v[2],
          A(0, 2) * v[0] + A(1, 2) * v[1] + A(2, 2) * v[2]};
}
template <class S>
M3<S> operator*(const M3<S>& A, const M3<S>& B) {
  M3<S> C;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C(i, j) = A(i, 0) * B(0, j) + A(i, 1) * B(1, j) + A(i, 2) * B(2, j);
  return C;
}
template <class S>
M3<S> transpose(const M3<S>& A) {
  M3<S> T;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) T(i, j) = A(j, i);
  return T;
}

template <class S>
struct SE3 {
  M3<S> R;
  V3<S> p;
  SE3() : R(M3<S>::identity()) {}
  SE3(const M3<S>& r, const V3<S>& t) : R(r), p(t) {}
  SE3 operator*(const SE3& o) const { return SE3(R * o.R, R * o.p + p); }
};
template <class S>
struct Motion {
  V3<S> lin, ang;
};
template <class S>
struct Force {
  V3<S> lin, ang;
};
template <class S>
Motion<S> operator+(const Motion<S>& a, const Motion<S>& b) { return {a.lin + b.lin, a.ang + b.ang}; }
template <class S>
Force<S> operator+(const Force<S>& a, const Force<S>& b) { return {a.lin + b.lin, a.ang + b.ang}; }
// SE3 actions (child -> parent: act ; parent -> child: actInv)
template <class S>
Motion<S> act(const SE3<S>& M, const Motion<S>& v) {
  const V3<S> w = M.R * v.ang;
  return {M.R * v.lin + cross(M.p, w), w};
}
template <class S>
Motion<S> actInv(const SE3<S>& M, const Motion<S>& v) { return {tmul(M.R, v.lin - cross(M.p, v.ang)), tmul(M.R, v.ang)}; }
template <class S>
Force<S> act(const SE3<S>& M, const Force<S>& f) {
  const V3<S> fl = M.R * f.lin;
  return {fl, M.R * f.ang + cross(M.p, fl)};
}
// motion x motion, motion x* force
template <class S>
Motion<S> cross(const Motion<S>& a, const Motion<S>& b) { return {cross(a.ang, b.lin) + cross(a.lin, b.ang), cross(a.ang, b.ang)}; }
template <class S>
Force<S> crossStar(const Motion<S>& a, const Force<S>& f) { return {cross(a.ang, f.lin), cross(a.ang, f.ang) + cross(a.lin, f.lin)}; }

struct BodyInertia {  // mass, com, rotational inertia about the com (body axes)
  double m = 0;
  double c[3] = {0, 0, 0};
  double I[9] = {0};
};
template <class S>
Force<S> inertiaTimes(const BodyInertia& Y, const Motion<S>& v) {
  const V3<S> c(S(Y.c[0]), S(Y.c[1]), S(Y.c[2]));
  const V3<S> fl = S(Y.m) * (v.lin - cross(c, v.ang));
  M3<S> I;
  for (int i = 0; i < 9; ++i) I.m[i] = S(Y.I[i]);
  return {fl, I * v.ang + cross(c, fl)};
}

struct RobotModel {
  int nj = 0;  // revolute joints; bodies 0..nj, body 0 = floating base
  std::vector<int> parent;
  std::vector<std::array<double, 9>> jointR;
  std::vector<std::array<double, 3>> jointP, axis;
  std::vector<BodyInertia> inertia;
  std::vector<double> qLower, qUpper;
  std::vector<int> frameBody;
  std::vector<std::array<double, 3>> frameP;
  double gravity = 9.81;
  // CentroidalModelInfo of the centroidal MPC (ocs2_centroidal_model/include/ocs2_centroidal_model/CentroidalModelInfo.h:47-70): model type
  // (0 FullCentroidalDynamics, 1 SingleRigidBodyDynamics) and the nominal inertia / com offset the latter uses; ignored by the whole-body OCP
  int centroidalModelType = 0;
  double inertiaNominal[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  double comToBaseNominal[3] = {0, 0, 0};
  double totalMass() const {
    double m = 0;
    for (const auto& b : inertia) m += b.m;
    return m;
  }
  int nv() const { return 6 + nj; }
};

// ---- joint models -----------------------------------------------------------------------------------------------
// SphericalZYX (Pinocchio joint-spherical-ZYX.hpp): R = Rz(q0) Ry(q1) Rx(q2); body-frame angular velocity w = S(q) qdot
template <class S>
void zyxRotation(const S* th, M3<S>& R, M3<S>& Sm) {
  const S c0 = cos(th[0]), s0 = sin(th[0]), c1 = cos(th[1]), s1 = sin(th[1]), c2 = cos(th[2]), s2 = sin(th[2]);
  R(0, 0) = c0 * c1;  R(0, 1) = c0 * s1 * s2 - s0 * c2;  R(0, 2) = c0 * s1 * c2 + s0 * s2;
  R(1, 0) = s0 * c1;  R(1, 1) = s0 * s1 * s2 + c0 * c2;  R(1, 2) = s0 * s1 * c2 - c0 * s2;
  R(2, 0) = -s1;      R(2, 1) = c1 * s2;                 R(2, 2) = c1 * c2;
  Sm(0, 0) = -s1;      Sm(0, 1) = S(0.0); Sm(0, 2) = S(1.0);
  Sm(1, 0) = c1 * s2;  Sm(1, 1) = c2;     Sm(1, 2) = S(0.0);
  Sm(2, 0) = c1 * c2;  Sm(2, 1) = -s2;    Sm(2, 2) = S(0.0);
}
// d/dt S(q) * qdot  (the SphericalZYX joint bias)
template <class S>
V3<S> zyxBias(const S* th, const S* thd) {
  const S c1 = cos(th[1]), s1 = sin(th[1]), c2 = cos(th[2]), s2 = sin(th[2]);
  return {-c1 * thd[1] * thd[0], (-s1 * s2 * thd[1] + c1 * c2 * thd[2]) * thd[0] - s2 * thd[2] * thd[1],
          (-s1 * c2 * thd[1] - c1 * s2 * thd[2]) * thd[0] - c2 * thd[2] * thd[1]};
}
template <class S>
M3<S> axisRotation(const double* a, const S& q) {  // Rodrigues, unit axis
  const S c = cos(q), s = sin(q), t = S(1.0) - c;
  M3<S> R;
  R(0, 0) = t * (a[0] * a[0]) + c;         R(0, 1) = t * (a[0] * a[1]) - s * a[2];  R(0, 2) = t * (a[0] * a[2]) + s * a[1];
  R(1, 0) = t * (a[0] * a[1]) + s * a[2];  R(1, 1) = t * (a[1] * a[1]) + c;         R(1, 2) = t * (a[1] * a[2]) - s * a[0];
  R(2, 0) = t * (a[0] * a[2]) - s * a[1];  R(2, 1) = t * (a[1] * a[2]) + s * a[0];  R(2, 2) = t * (a[2] * a[2]) + c;
  return R;
}

template <class S>
struct KinData {
  std::vector<SE3<S>> liMi, oMi;
  std::vector<Motion<S>> v, a;  // body-local spatial velocity / acceleration
  M3<S> Szyx;                   // SphericalZYX subspace of the base
};

// forwardKinematics(model, data, q, v, a) ; a0 = acceleration of the universe (0 for kinematics, -gravity for RNEA).
// qdd may be null (treated as zero).
template <class S>
void forwardKinematics(const RobotModel& m, const S* q, const S* qd, const S* qdd, const V3<S>& universeAcc, KinData<S>& d) {
  const int nb = m.nj + 1;
  d.liMi.resize(nb);
  d.oMi.resize(nb);
  d.v.resize(nb);
  d.a.resize(nb);
  // base: composite (Translation, SphericalZYX)
  M3<S> R0;
  zyxRotation(q + 3, R0, d.Szyx);
  d.liMi[0] = SE3<S>(R0, V3<S>(q[0], q[1], q[2]));
  d.oMi[0] = d.liMi[0];
  const V3<S> pd(qd[0], qd[1], qd[2]), thd(qd[3], qd[4], qd[5]);
  const V3<S> vl = tmul(R0, pd), wb = d.Szyx * thd;
  d.v[0] = {vl, wb};
  Motion<S> a0 = actInv(d.liMi[0], Motion<S>{universeAcc, V3<S>()});
  // joint bias c = (vl x wb, Sdot thd) ; S qdd term
  a0.lin = a0.lin + cross(vl, wb);
  a0.ang = a0.ang + zyxBias(q + 3, qd + 3);
  if (qdd) {
    a0.lin = a0.lin + tmul(R0, V3<S>(qdd[0], qdd[1], qdd[2]));
    a0.ang = a0.ang + d.Szyx * V3<S>(qdd[3], qdd[4], qdd[5]);
  }
  d.a[0] = a0;
  for (int i = 1; i < nb; ++i) {
    const int p = m.parent[i];
    M3<S> Rp;
    for (int k = 0; k < 9; ++k) Rp.m[k] = S(m.jointR[i][k]);
    const V3<S> pp(S(m.jointP[i][0]), S(m.jointP[i][1]), S(m.jointP[i][2]));
    d.liMi[i] = SE3<S>(Rp * axisRotation<S>(m.axis[i].data(), q[5 + i]), pp);
    d.oMi[i] = d.oMi[p] * d.liMi[i];
    const V3<S> ax(S(m.axis[i][0]), S(m.axis[i][1]), S(m.axis[i][2]));
    const Motion<S> vJ{V3<S>(), qd[5 + i] * ax};
    d.v[i] = actInv(d.liMi[i], d.v[p]) + vJ;
    Motion<S> ai = actInv(d.liMi[i], d.a[p]) + cross(d.v[i], vJ);
    if (qdd) ai.ang = ai.ang + qdd[5 + i] * ax;
    d.a[i] = ai;
  }
}

// rnea(q, v, a) (a may be null -> nonLinearEffects); tau has 6+nj entries
template <class S>
void rnea(const RobotModel& m, const S* q, const S* qd, const S* qdd, S* tau) {
  KinData<S> d;
  forwardKinematics(m, q, qd, qdd, V3<S>(S(0.0), S(0.0), S(m.gravity)), d);
  const int nb = m.nj + 1;
  std::vector<Force<S>> f(nb);
  for (int i = 0; i < nb; ++i) {
    const Force<S> h = inertiaTimes(m.inertia[i], d.v[i]);
    f[i] = inertiaTimes(m.inertia[i], d.a[i]) + crossStar(d.v[i], h);
  }
  for (int i = nb - 1; i >= 1; --i) {
    const V3<S> ax(S(m.axis[i][0]), S(m.axis[i][1]), S(m.axis[i][2]));
    tau[5 + i] = dot(ax, f[i].ang);
    f[m.parent[i]] = f[m.parent[i]] + act(d.liMi[i], f[i]);
  }
  // base: tau_lin = R f_lin (world-frame force), tau_ang = S^T n
  const V3<S> tl = d.liMi[0].R * f[0].lin, ta = tmul(d.Szyx, f[0].ang);
  for (int k = 0; k < 3; ++k) {
    tau[k] = tl[k];
    tau[3 + k] = ta[k];
  }
}

// 6x6 spatial inertia ([lin;ang] ordering), row-major
template <class S>
struct I6 {
  S m[36];
  I6() { for (auto& e : m) e = S(0.0); }
  S& operator()(int i, int j) { return m[6 * i + j]; }
  const S& operator()(int i, int j) const { return m[6 * i + j]; }
};
template <class S>
M3<S> skew(const V3<S>& v) {
  M3<S> K;
  K(0, 1) = -v[2]; K(0, 2) = v[1];
  K(1, 0) = v[2];  K(1, 2) = -v[0];
  K(2, 0) = -v[1]; K(2, 1) = v[0];
  return K;
}
template <class S>
I6<S> bodyInertia6(const BodyInertia& Y) {
  I6<S> I;
  const V3<S> c(S(Y.c[0]), S(Y.c[1]), S(Y.c[2]));
  const M3<S> C = skew(c);
  const M3<S> CC = C * C;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      I(i, j) = (i == j) ? S(Y.m) : S(0.0);
      I(i, 3 + j) = S(-Y.m) * C(i, j);
      I(3 + i, j) = S(Y.m) * C(i, j);
      I(3 + i, 3 + j) = S(Y.I[3 * i + j]) - S(Y.m) * CC(i, j);
    }
  return I;
}
// 6x6 motion transform child->parent of SE3 (R,p): [R, [p]x R; 0, R]
template <class S>
void motionXform(const SE3<S>& M, S X[36]) {
  const M3<S> pR = skew(M.p) * M.R;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      X[6 * i + j] = M.R(i, j);
      X[6 * i + 3 + j] = pR(i, j);
      X[6 * (3 + i) + j] = S(0.0);
      X[6 * (3 + i) + 3 + j] = M.R(i, j);
    }
}
// I_parent = X^-T I X^-1 with X = motion transform child->parent ; implemented via force transform F = X^-T = [R,0;[p]xR,R]
template <class S>
I6<S> transformInertia(const SE3<S>& M, const I6<S>& I) {
  S F[36];  // force transform child->parent
  const M3<S> pR = skew(M.p) * M.R;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      F[6 * i + j] = M.R(i, j);
      F[6 * i + 3 + j] = S(0.0);
      F[6 * (3 + i) + j] = pR(i, j);
      F[6 * (3 + i) + 3 + j] = M.R(i, j);
    }
  // I_p = F I F^T
  S T[36];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      S s(0.0);
      for (int k = 0; k < 6; ++k) s = s + F[6 * i + k] * I(k, j);
      T[6 * i + j] = s;
    }
  I6<S> O;
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      S s(0.0);
      for (int k = 0; k < 6; ++k) s = s + T[6 * i + k] * F[6 * j + k];
      O(i, j) = s;
    }
  return O;
}

// crba: full symmetric joint-space inertia M (nv x nv, row-major), composite-rigid-body algorithm
template <class S>
void crba(const RobotModel& m, const S* q, std::vector<S>& Mout) {
  const int nb = m.nj + 1, nv = m.nv();
  std::vector<S> zero(nv, S(0.0));
  KinData<S> d;
  forwardKinematics(m, q, zero.data(), static_cast<const S*>(nullptr), V3<S>(), d);
  std::vector<I6<S>> Y(nb);
  for (int i = 0; i < nb; ++i) Y[i] = bodyInertia6<S>(m.inertia[i]);
  Mout.assign(static_cast<size_t>(nv) * nv, S(0.0));
  // joint subspace columns in body-local coordinates
  auto subspace = [&](int body, int col, S s[6]) {
    for (int k = 0; k < 6; ++k) s[k] = S(0.0);
    if (body == 0) {
      if (col < 3) {
        for (int k = 0; k < 3; ++k) s[k] = d.liMi[0].R(col, k);  // R^T e_col
      } else {
        for (int k = 0; k < 3; ++k) s[3 + k] = d.Szyx(k, col - 3);
      }
    } else {
      for (int k = 0; k < 3; ++k) s[3 + k] = S(m.axis[body][k]);
    }
  };
  for (int i = nb - 1; i >= 0; --i) {
    const int ncol = (i == 0) ? 6 : 1, c0 = (i == 0) ? 0 : 5 + i;
    for (int c = 0; c < ncol; ++c) {
      S s[6], F[6];
      subspace(i, c, s);
      for (int r = 0; r < 6; ++r) {
        S acc(0.0);
        for (int k = 0; k < 6; ++k) acc = acc + Y[i](r, k) * s[k];
        F[r] = acc;
      }
      // diagonal block and propagation to ancestors
      Force<S> f{V3<S>(F[0], F[1], F[2]), V3<S>(F[3], F[4], F[5])};
      int j = i;
      while (true) {
        const int nc2 = (j == 0) ? 6 : 1, cj0 = (j == 0) ? 0 : 5 + j;
        for (int c2 = 0; c2 < nc2; ++c2) {
          S s2[6];
          subspace(j, c2, s2);
          const S val = s2[0] * f.lin[0] + s2[1] * f.lin[1] + s2[2] * f.lin[2] + s2[3] * f.ang[0] + s2[4] * f.ang[1] + s2[5] * f.ang[2];
          Mout[static_cast<size_t>(cj0 + c2) * nv + (c0 + c)] = val;
          Mout[static_cast<size_t>(c0 + c) * nv + (cj0 + c2)] = val;
        }
        if (j == 0) break;
        f = act(d.liMi[j], f);
        j = m.parent[j];
      }
    }
    if (i > 0) {
      const I6<S> Yp = transformInertia(d.liMi[i], Y[i]);
      for (int k = 0; k < 36; ++k) Y[m.parent[i]].m[k] = Y[m.parent[i]].m[k] + Yp.m[k];
    }
  }
}

// Frame quantities in LOCAL_WORLD_ALIGNED (origin at the frame, world axes)
template <class S>
struct FrameKin {
  V3<S> pos;
  M3<S> R;               // world rotation of the frame (= parent joint rotation; frames carry identity rotation)
  V3<S> vlin, vang;      // getFrameVelocity
  V3<S> alin, aang;      // getFrameClassicalAcceleration
};
template <class S>
FrameKin<S> frameKinematics(const RobotModel& m, const KinData<S>& d, int frame) {
  const int b = m.frameBody[frame];
  const V3<S> t(S(m.frameP[frame][0]), S(m.frameP[frame][1]), S(m.frameP[frame][2]));
  FrameKin<S> fk;
  fk.R = d.oMi[b].R;
  fk.pos = d.oMi[b].R * t + d.oMi[b].p;
  const SE3<S> iMf(M3<S>::identity(), t);
  const Motion<S> vf = actInv(iMf, d.v[b]);
  Motion<S> af = actInv(iMf, d.a[b]);
  af.lin = af.lin + cross(vf.ang, vf.lin);  // classical acceleration
  fk.vlin = fk.R * vf.lin;
  fk.vang = fk.R * vf.ang;
  fk.alin = fk.R * af.lin;
  fk.aang = fk.R * af.ang;
  return fk;
}

// Base (first six) columns of computeFrameJacobian(..., LOCAL_WORLD_ALIGNED) transposed times a wrench:
// J_b^T [f; m]  with f, m in world axes at the frame origin.
template <class S>
void baseJacobianTransposeWrench(const KinData<S>& d, const V3<S>& framePos, const V3<S>& f, const V3<S>& mo, S out[6]) {
  // translation columns: unit linear velocity in world -> J_lin = I, J_ang = 0
  for (int k = 0; k < 3; ++k) out[k] = f[k];
  // euler-rate columns: w_world = R0 S e_k ; v_frame = w x (p_f - p_b)
  const V3<S> r = framePos - d.oMi[0].p;
  for (int k = 0; k < 3; ++k) {
    const V3<S> w = d.oMi[0].R * V3<S>(d.Szyx(0, k), d.Szyx(1, k), d.Szyx(2, k));
    out[3 + k] = dot(cross(w, r), f) + dot(w, mo);
  }
}

}  // namespace orc
