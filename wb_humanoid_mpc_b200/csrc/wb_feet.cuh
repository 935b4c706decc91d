// K1/K3 building block: end-effector (foot) quantities and their tangents.
//
// Reference: PinocchioEndEffectorDynamicsCppAd (humanoid_nmpc/humanoid_wb_mpc/src/end_effector/PinocchioEndEffectorDynamicsCppAd.cpp:
// getPositionCppAd :241-255, getOrientationErrorWrtPlaneCppAd :497-517, getTwistCppAd :580-599, getAccelerationsCppAd :761-779), all in
// LOCAL_WORLD_ALIGNED, accelerations evaluated with a = computeGeneralizedAccelerations(x,u) and differentiated by CppAD.
//
// Here: the ankle body's spatial velocity / acceleration already sit in the dynamics workspace (pelvis coordinates).  Their
// tangents along a kinematic-chain direction are closed forms (dv_i/dq_k = S_k x (v_i - v_parent(k)), ...), the dependence on
// everything else enters only through qdd_b and is chained with the 6 x 93 matrix G of wb_dynamics.cuh.  The closed-form
// output map (rotate to world axes, classical acceleration, orientation error) is differentiated with a single-tangent dual.
#pragma once
#include "wb_dynamics.cuh"

namespace b200sqp {

constexpr int FQ = 18;       // foot quantities: pos, oriErr, vlin, vang, alin, aang
constexpr int FLOC = 36;     // local tangent directions per foot: pb(3) th(3) qleg(6) pd(3) thd(3) qdleg(6) aleg(6) qddb(6)

// map local direction index -> z index (0..92) or -1 for the qdd_b directions
HD int footLocalToZ(const WbDeviceModel& m, int c, int l) {
  if (l < 6) return l;                                            // pb, th
  if (l < 12) return 6 + m.legBody[c][l - 6] - 1;                 // q_leg
  if (l < 18) return NV + (l - 12);                               // pd, thd
  if (l < 24) return NV + 6 + m.legBody[c][l - 18] - 1;           // qd_leg
  if (l < 30) return NX + 12 + m.legBody[c][l - 24] - 1;          // qdd_leg (input)
  return -1;
}

// quaternionDistance(getQuaternionFromUnitVectors(a, e_z), Identity) = -(a x e_z)/norm  (ocs2_robotic_tools RotationTransforms.h:98-113,396-405)
HD void oriErrToPlane(DV3 a, D1* out) {
  const D1 cx = a.y, cy = -a.x;  // a x e_z = (a_y, -a_x, 0)
  const D1 w = D1{1.0, 0.0} + a.z;
  const D1 n = dsqrt(cx * cx + cy * cy + w * w);
  out[0] = -(cx / n);
  out[1] = -(cy / n);
  out[2] = D1{0.0, 0.0};
}

// Closed-form output map with one tangent.  Inputs are (value, tangent) pairs in pelvis coordinates.
HD void footOutputs(const double* Rb, const double* dRb, V3 pbase, V3 dpbase, V3 pf, V3 dpf, V3 zl, V3 dzl, V6 v, V6 dv, V6 a, V6 da,
                    D1* out /*18*/) {
  const DV3 P = dv3(pf, dpf), Z = dv3(zl, dzl);
  const DV3 vO = dv3(v.l, dv.l), om = dv3(v.a, dv.a), aO = dv3(a.l, da.l), al = dv3(a.a, da.a);
  const DV3 pw = dmv(Rb, dRb, P);
  out[0] = pw.x + D1{pbase.x, dpbase.x};
  out[1] = pw.y + D1{pbase.y, dpbase.y};
  out[2] = pw.z + D1{pbase.z, dpbase.z};
  oriErrToPlane(dmv(Rb, dRb, Z), out + 3);
  const DV3 vP = vO + dcross(om, P);
  const DV3 aP = aO + dcross(al, P) + dcross(om, vP);
  const DV3 vl = dmv(Rb, dRb, vP), va = dmv(Rb, dRb, om), al_ = dmv(Rb, dRb, aP), aa = dmv(Rb, dRb, al);
  out[6] = vl.x; out[7] = vl.y; out[8] = vl.z;
  out[9] = va.x; out[10] = va.y; out[11] = va.z;
  out[12] = al_.x; out[13] = al_.y; out[14] = al_.z;
  out[15] = aa.x; out[16] = aa.y; out[17] = aa.z;
}

// true spatial acceleration of body b (gravity removed, base acceleration added), pelvis coordinates
HD V6 trueAcc(const WbDeviceModel& m, const DynWs& w, int b) {
  V6 a = ld6(w.a[b]);
  const V3 gl = m.gravity * mk(w.Rb[6], w.Rb[7], w.Rb[8]);
  a.l = a.l - gl + mtv(w.Rb, ld3(w.qddb));
  a.a = a.a + mv(w.Sz, ld3(w.qddb + 3));
  return a;
}

// ---- foot values only (K3): FV[c][18] ---------------------------------------------------------------------------------------------
HD void footPhaseValues(Par P, const WbDeviceModel& m, const double* x, const DynWs& w, double* FV) {
  for (int c = P.tid; c < 2; c += P.nt) {
    const int b = m.frameBody[3 * c];
    const double zero9[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    const V3 z3 = mk(0, 0, 0);
    const V6 z6{z3, z3};
    D1 o[FQ];
    footOutputs(w.Rb, zero9, ld3(x), z3, ld3(w.pc[c]), z3, mk(w.R[b][2], w.R[b][5], w.R[b][8]), z3, ld6(w.v[b]), z6, trueAcc(m, w, b), z6, o);
    for (int k = 0; k < FQ; ++k) FV[FQ * c + k] = o[k].v;
  }
}

// ---- foot values + local tangents: items (c, l) ; JFl[c][k + FQ*l], FV[c][k] -----------------------------------------------------------
// Item layout: three groups padded to one warp each, so that a warp mixes at most three kinds of direction:
//   items 0-23 leg joint positions / velocities, 32-55 leg joint accelerations / base-acceleration directions, 64-87 base pose / base twist.
HD void footPhaseLocalTangents(Par P, const WbDeviceModel& m, const double* x, const DynWs& w, double* FV, double* JFl) {
  for (int it = P.tid; it < 96; it += P.nt) {
    const int grp = it >> 5, q = it & 31;
    if (q >= 24) continue;
    const int c = q & 1, h = q >> 1;   // h in 0..11
    const int l = (grp == 0) ? (h < 6 ? 6 + h : 12 + h) : ((grp == 1) ? 24 + h : (h < 6 ? h : 6 + h));
    const int b = m.frameBody[3 * c];
    const V3 z3 = mk(0, 0, 0);
    const double zero9[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    const V3 pf = ld3(w.pc[c]), zl = mk(w.R[b][2], w.R[b][5], w.R[b][8]);
    const V6 v = ld6(w.v[b]), ar = ld6(w.a[b]), a = trueAcc(m, w, b);
    const V6 v0 = ld6(w.v0);
    V3 dpb = z3, dpf = z3, dzl = z3;
    V6 dv{z3, z3}, da{z3, z3};
    const double* dRb = zero9;
    if (l < 3) {
      dpb = mk(l == 0, l == 1, l == 2);
    } else if (l < 6) {
      const int k = l - 3;
      dRb = w.dRb[k];
      dv = ld6(w.dv0[k]);
      da = ld6(w.da0[k]) + mcross(dv, v - v0);
      // remove the gravity term's tangent, add the base-acceleration terms' tangents
      da.l = da.l - m.gravity * mk(dRb[6], dRb[7], dRb[8]) + mtv(dRb, ld3(w.qddb));
      da.a = da.a + mv(w.dSz[k], ld3(w.qddb + 3));
    } else if (l < 12) {
      const int k = m.legBody[c][l - 6], pk = m.parent[k];
      const V6 S = ld6(w.S[k]);
      dpf = cross(S.a, pf - ld3(w.p[k]));
      dzl = cross(S.a, zl);
      dv = mcross(S, v - ld6(w.v[pk]));
      da = mcross(S, ar - ld6(w.a[pk])) + mcross(ld6(w.psd[k]), v - ld6(w.v[pk]));
    } else if (l < 18) {
      const int dir = 3 + (l - 12);
      dv = ld6(w.dv0[dir]);
      da = ld6(w.da0[dir]) + mcross(dv, v - v0);
    } else if (l < 24) {
      const int k = m.legBody[c][l - 18];
      const V6 S = ld6(w.S[k]);
      dv = S;
      da = 2.0 * ld6(w.psd[k]) + mcross(S, v);
    } else if (l < 30) {
      da = ld6(w.S[m.legBody[c][l - 24]]);
    } else {
      const int j = l - 30;
      const V3 e = mk(j % 3 == 0, j % 3 == 1, j % 3 == 2);
      if (j < 3) da.l = mtv(w.Rb, e);
      else da.a = mv(w.Sz, e);
    }
    D1 o[FQ];
    footOutputs(w.Rb, dRb, ld3(x), dpb, pf, dpf, zl, dzl, v, dv, a, da, o);
    for (int k = 0; k < FQ; ++k) JFl[(c * FLOC + l) * FQ + k] = o[k].d;
    if (l == 0)
      for (int k = 0; k < FQ; ++k) FV[FQ * c + k] = o[k].v;
  }
}

// all operational-frame positions in world coordinates (value-only: FP[f][3]) and, optionally, their tangents w.r.t. th(3) and the
// 12 leg joints: DFP[f][15][3]  (base translation cancels in all pairwise distances)
template <bool DERIV>
HD void framePhasePositions(Par P, const WbDeviceModel& m, const double* x, const DynWs& w, double* FP, double* DFP) {
  for (int it = P.tid; it < NFRAMES * (DERIV ? 16 : 1); it += P.nt) {
    const int f = it % NFRAMES, l = it / NFRAMES;  // l = 0: value ; l = 1..15: tangent direction l-1
    const int b = m.frameBody[f];
    const V3 pl = ld3(w.p[b]) + mv(w.R[b], ld3(m.frameP[f]));
    if (l == 0) {
      st3(FP + 3 * f, ld3(x) + mv(w.Rb, pl));
    } else {
      const int dl = l - 1;
      V3 t = mk(0, 0, 0);
      if (dl < 3) {
        t = mv(w.dRb[dl], pl);
      } else {
        const int k = dl - 3 + 1;  // leg joints are bodies 1..12
        if (m.subtree[k] >> b & 1u) t = mv(w.Rb, cross(ld3(w.S[k] + 3), pl - ld3(w.p[k])));
      }
      st3(DFP + (f * 15 + dl) * 3, t);
    }
  }
}

}  // namespace b200sqp
