// ORACLE (test infrastructure only; never linked into libb200sqp.so) -- centroidal flow map of the humanoid centroidal MPC.
//
// Restates, templated on the scalar like the reference (CppAD there, Dual<N> here):
//   PinocchioCentroidalDynamicsAD::getValueCppAd        lib/ocs2_ros2/ocs2_pinocchio/ocs2_centroidal_model/src/PinocchioCentroidalDynamicsAD.cpp:75-94
//   updateCentroidalDynamics (FullCentroidalDynamics)    .../src/ModelHelperFunctions.cpp:46-58   -> pinocchio::computeCentroidalMap: Ag, com
//   getPositionComToContactPointInWorldFrame             .../src/ModelHelperFunctions.cpp:139-145
//   getNormalizedCentroidalMomentumRate                  .../src/ModelHelperFunctions.cpp:167-194
//   CentroidalModelPinocchioMapping::getPinocchioJointVelocity   .../src/CentroidalModelPinocchioMapping.cpp:84-107
//   computeFloatingBaseCentroidalMomentumMatrixInverse   .../include/ocs2_centroidal_model/implementation/ModelHelperFunctionsImpl.h:40-47
// Layout (CentroidalModelInfo, two 6-DoF contacts, AccessHelperFunctionsImpl.h):
//   x = [ normalized momentum h/m (lin 3, ang 3) ; q = (base position 3, Euler ZYX 3, joints nj) ],  u = [ wrench_l (f 3, tau 3) ; wrench_r ; qdot_j nj ]
// Pinocchio (an un-vendored dependency) supplies computeCentroidalMap; its published definition is restated: column k of Ag is the
// spatial momentum of the whole robot, about the centre of mass in world-aligned axes, for unit generalized velocity e_k (base joint =
// Composite(Translation, SphericalZYX): v_base = world translation velocity, Euler-angle rates).
#pragma once
#include "rbd.hpp"

namespace orc {

template <class S>
struct CentroidalData {
  std::vector<S> Ag;  // 6 x nv, row-major
  V3<S> com;
  KinData<S> kin;
};

// total momentum (about the world origin, world axes) and centre of mass for generalized velocity qd
template <class S>
Force<S> worldMomentum(const RobotModel& m, const KinData<S>& d) {
  Force<S> h{V3<S>(), V3<S>()};
  for (int i = 0; i <= m.nj; ++i) h = h + act(d.oMi[i], inertiaTimes(m.inertia[i], d.v[i]));
  return h;
}
template <class S>
V3<S> centerOfMass(const RobotModel& m, const KinData<S>& d) {
  V3<S> s;
  for (int i = 0; i <= m.nj; ++i) {
    const V3<S> c(S(m.inertia[i].c[0]), S(m.inertia[i].c[1]), S(m.inertia[i].c[2]));
    s = s + S(m.inertia[i].m) * (d.oMi[i].R * c + d.oMi[i].p);
  }
  return (S(1.0) / S(m.totalMass())) * s;
}

// pinocchio::computeCentroidalMap(model, data, q): data.Ag (6 x nv) and data.com[0]
template <class S>
void computeCentroidalMap(const RobotModel& m, const S* q, CentroidalData<S>& c) {
  const int nv = m.nv();
  c.Ag.assign(6 * nv, S(0.0));
  std::vector<S> qd(nv, S(0.0));
  if (m.centroidalModelType == 1) {
    // updateCentroidalDynamics, SingleRigidBodyDynamics branch (ModelHelperFunctions.cpp:61-79): Ag = [Ab 0] from the nominal inertia and
    // com offset rotated with the base; com = base position - R r_nominal; the frame placements still follow the full kinematics
    forwardKinematics<S>(m, q, qd.data(), static_cast<const S*>(nullptr), V3<S>(), c.kin);
    const M3<S>& R = c.kin.oMi[0].R;
    const M3<S> T = R * c.kin.Szyx;   // Euler-rate -> global angular velocity
    const V3<S> rw = R * V3<S>(S(m.comToBaseNominal[0]), S(m.comToBaseNominal[1]), S(m.comToBaseNominal[2]));
    M3<S> In;
    for (int k = 0; k < 9; ++k) In.m[k] = S(m.inertiaNominal[k]);
    const M3<S> A12 = skew(rw) * T, A22 = (R * In) * (transpose(R) * T);
    const S mass(m.totalMass());
    for (int r = 0; r < 3; ++r) {
      c.Ag[r * nv + r] = mass;
      for (int k = 0; k < 3; ++k) {
        c.Ag[r * nv + 3 + k] = mass * A12(r, k);
        c.Ag[(3 + r) * nv + 3 + k] = A22(r, k);
      }
    }
    c.com = V3<S>(q[0] - rw[0], q[1] - rw[1], q[2] - rw[2]);
    return;
  }
  for (int k = 0; k < nv; ++k) {
    qd[k] = S(1.0);
    forwardKinematics<S>(m, q, qd.data(), static_cast<const S*>(nullptr), V3<S>(), c.kin);
    if (k == 0) c.com = centerOfMass(m, c.kin);
    const Force<S> h = worldMomentum(m, c.kin);
    const V3<S> ang = h.ang - cross(c.com, h.lin);  // shift the reference point from the world origin to the centre of mass
    c.Ag[0 * nv + k] = h.lin[0];
    c.Ag[1 * nv + k] = h.lin[1];
    c.Ag[2 * nv + k] = h.lin[2];
    c.Ag[3 * nv + k] = ang[0];
    c.Ag[4 * nv + k] = ang[1];
    c.Ag[5 * nv + k] = ang[2];
    qd[k] = S(0.0);
  }
}

template <class S>
void inv3(const S* A, S* Ai) {  // row-major 3x3
  const S c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
  const S id = S(1.0) / (A[0] * c00 + A[1] * c01 + A[2] * c02);
  Ai[0] = c00 * id; Ai[1] = (A[2] * A[7] - A[1] * A[8]) * id; Ai[2] = (A[1] * A[5] - A[2] * A[4]) * id;
  Ai[3] = c01 * id; Ai[4] = (A[0] * A[8] - A[2] * A[6]) * id; Ai[5] = (A[2] * A[3] - A[0] * A[5]) * id;
  Ai[6] = c02 * id; Ai[7] = (A[1] * A[6] - A[0] * A[7]) * id; Ai[8] = (A[0] * A[4] - A[1] * A[3]) * id;
}

// xdot = f(x, u); contactFrame[c] = index of the contact frame of foot c in model.frameBody / frameP
template <class S>
void centroidalFlowMap(const RobotModel& m, const int contactFrame[2], const S* x, const S* u, S* xdot) {
  const int nj = m.nj, nv = 6 + nj;
  const S* q = x + 6;
  CentroidalData<S> c;
  computeCentroidalMap(m, q, c);
  const S mass(m.totalMass());
  // getNormalizedCentroidalMomentumRate
  V3<S> lin(S(0.0), S(0.0), S(-9.81) * mass), ang;
  for (int k = 0; k < 2; ++k) {
    const int f = contactFrame[k], b = m.frameBody[f];
    const V3<S> fp(S(m.frameP[f][0]), S(m.frameP[f][1]), S(m.frameP[f][2]));
    const V3<S> r = c.kin.oMi[b].R * fp + c.kin.oMi[b].p - c.com;   // positionComToContactPointInWorldFrame
    const V3<S> F(u[6 * k], u[6 * k + 1], u[6 * k + 2]), T(u[6 * k + 3], u[6 * k + 4], u[6 * k + 5]);
    lin = lin + F;
    ang = ang + cross(r, F) + T;
  }
  const S im = S(1.0) / mass;
  xdot[0] = im * lin[0]; xdot[1] = im * lin[1]; xdot[2] = im * lin[2];
  xdot[3] = im * ang[0]; xdot[4] = im * ang[1]; xdot[5] = im * ang[2];
  // getPinocchioJointVelocity: v_b = Ab^-1 (m hbar - Aj qdot_j) with the block inverse of computeFloatingBaseCentroidalMomentumMatrixInverse
  S mom[6];
  for (int r = 0; r < 6; ++r) {
    mom[r] = mass * x[r];
    for (int j = 0; j < nj; ++j) mom[r] = mom[r] - c.Ag[r * nv + 6 + j] * u[12 + j];
  }
  S Ab22[9], Ab22i[9], Ab12[9];
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 3; ++k) {
      Ab22[3 * r + k] = c.Ag[(3 + r) * nv + 3 + k];
      Ab12[3 * r + k] = c.Ag[r * nv + 3 + k];
    }
  inv3(Ab22, Ab22i);
  const S mA = c.Ag[0];  // mass = Ab(0, 0)
  S wb[3];
  for (int r = 0; r < 3; ++r) wb[r] = Ab22i[3 * r] * mom[3] + Ab22i[3 * r + 1] * mom[4] + Ab22i[3 * r + 2] * mom[5];
  for (int r = 0; r < 3; ++r) {
    const S t = Ab12[3 * r] * wb[0] + Ab12[3 * r + 1] * wb[1] + Ab12[3 * r + 2] * wb[2];
    xdot[6 + r] = (S(1.0) / mA) * mom[r] - (S(1.0) / mA) * t;
    xdot[9 + r] = wb[r];
  }
  for (int j = 0; j < nj; ++j) xdot[12 + j] = u[12 + j];
}

}  // namespace orc
