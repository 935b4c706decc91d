// Joint-torque map of the MRT controllers: planned (x, u) -> feed-forward joint torques.
//
// Reference: computeJointTorques (humanoid_nmpc/humanoid_common_mpc/src/pinocchio_model/DynamicsHelperFunctions.cpp:232-270, called per control
// tick by WBMpcMrtJointController.cpp:141): crba + nonLinearEffects + the two contact-frame Jacobians, the base acceleration from
// computeBaseAcceleration (block-diagonal M_bb inverse, as in the flow map) and
//   tau_j = M_j [qdd_b ; qdd_j] + nle_j - (J_l' W_l + J_r' W_r)_j .
// B200 formulation: value-only, so one THREAD per (instance, node) runs a world-frame Newton-Euler recursion twice (zero base acceleration
// -> base wrench residual -> qdd_b -> torques); a batch of planned trajectories maps to a plain grid, no shared memory, no barriers.
#pragma once
#include "wb_model.cuh"

namespace b200sqp {

struct TqBody {
  double R[9];
  V3 p, c, w, al, a;   // origin, com, angular velocity, angular acceleration, origin acceleration (world axes)
  V3 ax;               // joint axis in world axes
};

HD void zyxRS(const double* th, double* R, double* S) {   // R = Rz Ry Rx ; body-frame angular velocity = S thd (SphericalZYX)
  const double c0 = cos(th[0]), s0 = sin(th[0]), c1 = cos(th[1]), s1 = sin(th[1]), c2 = cos(th[2]), s2 = sin(th[2]);
  R[0] = c0 * c1; R[1] = c0 * s1 * s2 - s0 * c2; R[2] = c0 * s1 * c2 + s0 * s2;
  R[3] = s0 * c1; R[4] = s0 * s1 * s2 + c0 * c2; R[5] = s0 * s1 * c2 - c0 * s2;
  R[6] = -s1;     R[7] = c1 * s2;                R[8] = c1 * c2;
  S[0] = -s1;     S[1] = 0.0; S[2] = 1.0;
  S[3] = c1 * s2; S[4] = c2;  S[5] = 0.0;
  S[6] = c1 * c2; S[7] = -s2; S[8] = 0.0;
}

// generalized forces of the tree for generalized accelerations (pdd, thdd, qdd_j) minus the contact wrenches: out[0..6) base, out[6..) joints.
// Also returns M_lin = mtot and the base-frame composite rotational inertia about the base origin (for the block-diagonal M_bb inverse).
HD void tqInverseDynamics(const WbDeviceModel& m, const double* x, const double* u, const double* qddb, TqBody* B, double* out, double* IcBase) {
  const double* q = x;
  const double* qd = x + NV;
  double S[9];
  zyxRS(q + 3, B[0].R, S);
  B[0].p = ld3(q);
  const V3 thd = ld3(qd + 3);
  const V3 wb = mv(S, thd);   // body frame
  B[0].w = mv(B[0].R, wb);
  // SphericalZYX bias: d/dt S * thd
  {
    const double c1 = cos(q[4]), s1 = sin(q[4]), c2 = cos(q[5]), s2 = sin(q[5]);
    const double d1 = qd[4], d2 = qd[5], d0 = qd[3];
    const V3 bias = mk(-c1 * d1 * d0, (-s1 * s2 * d1 + c1 * c2 * d2) * d0 - s2 * d2 * d1, (-s1 * c2 * d1 - c1 * s2 * d2) * d0 - c2 * d2 * d1);
    const V3 sdd = qddb ? mv(S, ld3(qddb + 3)) : mk(0, 0, 0);
    B[0].al = mv(B[0].R, bias + sdd);
  }
  B[0].a = (qddb ? ld3(qddb) : mk(0, 0, 0)) + mk(0, 0, m.gravity);   // gravity as an upward acceleration of the base
  B[0].ax = mk(0, 0, 0);
  for (int i = 1; i < NB; ++i) {
    const int pa = m.parent[i];
    const double* a = m.axis[i];
    const double th = q[5 + i], c = cos(th), s = sin(th), t = 1.0 - c;
    const double Rq[9] = {t * a[0] * a[0] + c,        t * a[0] * a[1] - s * a[2], t * a[0] * a[2] + s * a[1],
                          t * a[0] * a[1] + s * a[2], t * a[1] * a[1] + c,        t * a[1] * a[2] - s * a[0],
                          t * a[0] * a[2] - s * a[1], t * a[1] * a[2] + s * a[0], t * a[2] * a[2] + c};
    double Rj[9];
    mm3(m.jR[i], Rq, Rj);
    mm3(B[pa].R, Rj, B[i].R);
    const V3 r = mv(B[pa].R, ld3(m.jp[i]));
    B[i].p = B[pa].p + r;
    B[i].ax = mv(B[i].R, ld3(a));
    const double qdi = qd[5 + i], qddi = u[12 + (i - 1)];
    B[i].w = B[pa].w + qdi * B[i].ax;
    B[i].al = B[pa].al + qddi * B[i].ax + qdi * cross(B[pa].w, B[i].ax);
    B[i].a = B[pa].a + cross(B[pa].al, r) + cross(B[pa].w, cross(B[pa].w, r));
  }
  // body wrenches about the world origin, accumulated over subtrees (children are numbered after their parents)
  V3 F[NB], Mo[NB];
  double Ic[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < NB; ++i) {
    const V3 rc = mv(B[i].R, ld3(m.com[i]));
    B[i].c = B[i].p + rc;
    const V3 ac = B[i].a + cross(B[i].al, rc) + cross(B[i].w, cross(B[i].w, rc));
    double RI[9], Rt[9], Iw[9];
    mm3(B[i].R, m.Icom[i], RI);
    for (int r = 0; r < 3; ++r)
      for (int k = 0; k < 3; ++k) Rt[3 * r + k] = B[i].R[3 * k + r];
    mm3(RI, Rt, Iw);
    F[i] = m.mass[i] * ac;
    const V3 N = mv(Iw, B[i].al) + cross(B[i].w, mv(Iw, B[i].w));
    Mo[i] = N + cross(B[i].c, F[i]);
    if (IcBase) {   // composite rotational inertia about the base origin, world axes
      const V3 d = B[i].c - B[0].p;
      const double dd = dot(d, d), dv[3] = {d.x, d.y, d.z};
      for (int r = 0; r < 3; ++r)
        for (int k = 0; k < 3; ++k) Ic[3 * r + k] += Iw[3 * r + k] + m.mass[i] * ((r == k ? dd : 0.0) - dv[r] * dv[k]);
    }
  }
  for (int c = 0; c < 2; ++c) {   // contact wrenches act on the foot bodies (LOCAL_WORLD_ALIGNED at the contact frame)
    const int b = m.frameBody[3 * c];
    const V3 pf = B[b].p + mv(B[b].R, ld3(m.frameP[3 * c]));
    const V3 Fe = ld3(u + 6 * c), Me = ld3(u + 6 * c + 3);
    F[b] = F[b] - Fe;
    Mo[b] = Mo[b] - (Me + cross(pf, Fe));
  }
  for (int i = NB - 1; i >= 1; --i) {
    out[5 + i] = dot(B[i].ax, Mo[i] - cross(B[i].p, F[i]));
    F[m.parent[i]] = F[m.parent[i]] + F[i];
    Mo[m.parent[i]] = Mo[m.parent[i]] + Mo[i];
  }
  st3(out, F[0]);
  const V3 nb = mtv(B[0].R, Mo[0] - cross(B[0].p, F[0]));   // base frame
  st3(out + 3, mtv(S, nb));
  if (IcBase) {   // M_ang = S' (R' Ic R) S
    double Rt[9], T1[9], T2[9], St[9];
    for (int r = 0; r < 3; ++r)
      for (int k = 0; k < 3; ++k) {
        Rt[3 * r + k] = B[0].R[3 * k + r];
        St[3 * r + k] = S[3 * k + r];
      }
    mm3(Rt, Ic, T1);
    mm3(T1, B[0].R, T2);
    mm3(T2, S, T1);
    mm3(St, T1, IcBase);
  }
}

HD void inv3x3(const double* A, double* Ai) {   // cofactor inverse (Eigen's 3x3 inverse)
  const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
  const double id = 1.0 / (A[0] * c00 + A[1] * c01 + A[2] * c02);
  Ai[0] = c00 * id; Ai[1] = (A[2] * A[7] - A[1] * A[8]) * id; Ai[2] = (A[1] * A[5] - A[2] * A[4]) * id;
  Ai[3] = c01 * id; Ai[4] = (A[0] * A[8] - A[2] * A[6]) * id; Ai[5] = (A[2] * A[3] - A[0] * A[5]) * id;
  Ai[6] = c02 * id; Ai[7] = (A[1] * A[6] - A[0] * A[7]) * id; Ai[8] = (A[0] * A[4] - A[1] * A[3]) * id;
}

// tau [NJ] and, optionally, the base acceleration [6] of computeBaseAcceleration
HD void wbJointTorques(const WbDeviceModel& m, const double* x, const double* u, double* tau, double* qddbOut) {
  TqBody B[NB];
  double g[NV], Mang[9], Mai[9], qddb[6];
  tqInverseDynamics(m, x, u, nullptr, B, g, Mang);   // generalized forces at zero base acceleration, contact wrenches included
  inv3x3(Mang, Mai);
  for (int k = 0; k < 3; ++k) {
    qddb[k] = -g[k] / m.mtot;
    qddb[3 + k] = -(Mai[3 * k] * g[3] + Mai[3 * k + 1] * g[4] + Mai[3 * k + 2] * g[5]);
  }
  tqInverseDynamics(m, x, u, qddb, B, g, nullptr);
  for (int j = 0; j < NJ; ++j) tau[j] = g[6 + j];
  if (qddbOut)
    for (int k = 0; k < 6; ++k) qddbOut[k] = qddb[k];
}

#ifdef __CUDACC__
// x [count][58], u [count][35] -> tau [count][23], qddb [count][6] (optional)
__global__ void __launch_bounds__(128) wb_torque_kernel(const WbDeviceModel* model, int count, const double* __restrict__ x, const double* __restrict__ u,
                                                        double* __restrict__ tau, double* __restrict__ qddb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  wbJointTorques(*model, x + static_cast<size_t>(i) * NX, u + static_cast<size_t>(i) * NU, tau + static_cast<size_t>(i) * NJ,
                 qddb ? qddb + static_cast<size_t>(i) * 6 : nullptr);
}
#endif

}  // namespace b200sqp
