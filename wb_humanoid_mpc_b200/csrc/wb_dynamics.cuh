// K1/K3 building block: whole-body flow map xdot = f(x,u) and the analytic Jacobian of the base acceleration.
//
// Reference path: WBAccelDynamicsAD::systemFlowMap -> computeStateDerivative -> computeBaseAcceleration
//   (humanoid_nmpc/humanoid_wb_mpc/src/dynamics/DynamicsHelperFunctions.cpp:51-134) with the block-diagonal base inertia
//   inverse of humanoid_common_mpc/src/pinocchio_model/DynamicsHelperFunctions.cpp:196-218; the reference differentiates it
//   with CppAD (ocs2_core/src/automatic_differentation/CppAdInterface.cpp:165-187).
//
// B200 formulation (not a port of Pinocchio/CppAD):
//   * everything is expressed in pelvis coordinates about the pelvis origin, so the base pose enters only through
//     v0 = (R'pdot, S thdot), a0 = (vl x wb + R'g, Sdot thdot) and the final rotations q̈_lin = R N_lin/m, q̈_ang = S^-1 Ic^-1 N_ang;
//   * M_bj qdd_j + nle_b is one Newton-Euler sweep; composite quantities in a common frame are plain sums over subtrees;
//   * d(total wrench)/d(q_k, qd_k, qdd_k) use the closed forms for single-DoF joints
//       dF/dq_k = S_k x* fC_k + IC_k psidd_k + 2 BC_k psid_k,  dF/dqd_k = 2 IC_k psid_k + 2 BC_k S_k,  dF/dqdd_k = IC_k S_k
//     (Singh, Russell, Wensing, "Efficient analytical derivatives of rigid-body dynamics using spatial vector algebra", RA-L 2022),
//     so one tangent direction is O(1) work once the composites exist: one thread per column of the 6 x 93 Jacobian.
// All functions are phase functions: every work item writes only its own outputs, phases are separated by a block barrier.
#pragma once
#include "dense_par.cuh"
#include "wb_model.cuh"

namespace b200sqp {

// shared-memory workspace of one evaluation point (doubles)
struct DynWs {
  // ---- what the value-only path (K3) needs: the first DYN_VALUE_DOUBLES doubles ------------------------------------------------------
  // base
  double Rb[9], Sz[9], SzInv[9], v0[6], a0[6];
  // bodies (pelvis coordinates)
  double Rj[NB][9];  // joint placement times joint rotation (parent-relative), computed body-parallel before the chain sweep
  double R[NB][9], p[NB][3], S[NB][6], v[NB][6], a[NB][6];
  double f[NB][6];   // per body; folded IN PLACE into subtree composites by dynPhaseComposite
  // finals
  double pc[2][3], lf[2][3], lm[2][3];  // contact points, local contact force / moment
  double N[6], IcInv[9], y[3], qddb[6];
  // Spatial inertia per body, folded in place into subtree composites.  The value-only path keeps just the rotational 3 x 3 block of every
  // body, packed as [NB][9] at the start of this array (it needs the root composite's block only): its workspace ends after NB * 9 doubles.
  double I[NB][36];
  // ---- derivative path (K1a) only -----------------------------------------------------------------------------------------------------
  double psd[NB][6], psdd[NB][6], Bm[NB][36];
  double dv0[9][6], da0[9][6], dRb[3][9], dSz[3][9];  // tangents w.r.t. th(3), pdot(3), thdot(3)
};
// doubles of the value-only prefix of DynWs (through the packed [NB][9] inertia blocks)
constexpr int DYN_VALUE_DOUBLES = 9 + 9 + 9 + 6 + 6 + NB * (9 + 9 + 3 + 6 + 6 + 6 + 6) + 18 + 6 + 9 + 3 + 6 + NB * 9;
static_assert(offsetof(DynWs, I) + NB * 9 * sizeof(double) == DYN_VALUE_DOUBLES * sizeof(double), "value-only prefix of DynWs");


// Base kinematics with one tangent direction (dir in 0..8 = th, pdot, thdot; dir < 0: values only).  Straight-line code: the value lane and
// the nine tangent lanes of a warp run the same instruction stream.
HD void baseKinematics(const double* x, int dir, double g, D1* R, D1* S, D1* v0, D1* a0) {
  D1 th[3], pd[3], td[3], sn[3], cs[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    th[k] = D1{x[3 + k], dir == k ? 1.0 : 0.0};
    pd[k] = D1{x[NV + k], dir == 3 + k ? 1.0 : 0.0};
    td[k] = D1{x[NV + 3 + k], dir == 6 + k ? 1.0 : 0.0};
    double sv, cv;
    sincos(th[k].v, &sv, &cv);
    sn[k] = D1{sv, cv * th[k].d};
    cs[k] = D1{cv, -sv * th[k].d};
  }
  // Rz(th0) Ry(th1) Rx(th2) and the body-frame angular-velocity map of Pinocchio's JointModelSphericalZYX
  const D1 c0 = cs[0], s0 = sn[0], c1 = cs[1], s1 = sn[1], c2 = cs[2], s2 = sn[2];
  const D1 zero{0.0, 0.0}, one{1.0, 0.0};
  R[0] = c0 * c1; R[1] = c0 * s1 * s2 - s0 * c2; R[2] = c0 * s1 * c2 + s0 * s2;
  R[3] = s0 * c1; R[4] = s0 * s1 * s2 + c0 * c2; R[5] = s0 * s1 * c2 - c0 * s2;
  R[6] = -s1;     R[7] = c1 * s2;                R[8] = c1 * c2;
  S[0] = -s1;     S[1] = zero; S[2] = one;
  S[3] = c1 * s2; S[4] = c2;   S[5] = zero;
  S[6] = c1 * c2; S[7] = -s2;  S[8] = zero;
  // vl = R' pdot, wb = S thdot
  const DV3 vl{R[0] * pd[0] + R[3] * pd[1] + R[6] * pd[2], R[1] * pd[0] + R[4] * pd[1] + R[7] * pd[2], R[2] * pd[0] + R[5] * pd[1] + R[8] * pd[2]};
  const DV3 wb{S[0] * td[0] + S[1] * td[1] + S[2] * td[2], S[3] * td[0] + S[4] * td[1] + S[5] * td[2], S[6] * td[0] + S[7] * td[1] + S[8] * td[2]};
  v0[0] = vl.x; v0[1] = vl.y; v0[2] = vl.z; v0[3] = wb.x; v0[4] = wb.y; v0[5] = wb.z;
  // a0 = (vl x wb + R' g e_z, Sdot thdot)
  const DV3 c = dcross(vl, wb);
  a0[0] = c.x + g * R[6];
  a0[1] = c.y + g * R[7];
  a0[2] = c.z + g * R[8];
  a0[3] = -(c1 * td[1] * td[0]);
  a0[4] = (c1 * c2 * td[2] - s1 * s2 * td[1]) * td[0] - s2 * td[2] * td[1];
  a0[5] = -((s1 * c2 * td[1] + c1 * s2 * td[2]) * td[0]) - c2 * td[2] * td[1];
}

// (Ir - m c^ c^) w : rotational block of the spatial inertia about the pelvis origin applied to w
HD V3 mvIr(const double* Ir, double ms, V3 c, V3 w) { return mv(Ir, w) - ms * cross(c, cross(c, w)); }
HD void inv3(const double* A, double* Ai) {
  const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
  const double det = A[0] * c00 + A[1] * c01 + A[2] * c02, id = 1.0 / det;
  Ai[0] = c00 * id; Ai[1] = (A[2] * A[7] - A[1] * A[8]) * id; Ai[2] = (A[1] * A[5] - A[2] * A[4]) * id;
  Ai[3] = c01 * id; Ai[4] = (A[0] * A[8] - A[2] * A[6]) * id; Ai[5] = (A[2] * A[3] - A[0] * A[5]) * id;
  Ai[6] = c02 * id; Ai[7] = (A[1] * A[6] - A[0] * A[7]) * id; Ai[8] = (A[0] * A[4] - A[1] * A[3]) * id;
}

// ---- phase 1a: base quantities (warp 0: lane 0 values, lanes 1..9 tangents) and joint transforms (warp 1), all independent ---------
template <bool DERIV>
HD void dynPhaseJoints(Par P, const WbDeviceModel& m, const double* x, DynWs& w) {
  // item layout: the base-kinematics items (value + nine tangent directions) fill the first warp of the derivative path and the joints start on
  // the next one; the value-only path (one base item, 23 joints) packs everything into 24 items -- a single round of the one-warp K3
  constexpr int J0 = DERIV ? 32 : 1;
  for (int it = P.tid; it < J0 + NJ; it += P.nt) {
    if (it < (DERIV ? 10 : 1)) {
      const int dir = it - 1;
      D1 R[9], S[9], v0[6], a0[6];
      baseKinematics(x, dir, m.gravity, R, S, v0, a0);
      if (dir < 0) {
        double Sv[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          w.Rb[k] = R[k].v;
          w.Sz[k] = Sv[k] = S[k].v;
        }
        inv3(Sv, w.SzInv);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          w.v0[k] = v0[k].v;
          w.a0[k] = a0[k].v;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          w.dv0[dir][k] = v0[k].d;
          w.da0[dir][k] = a0[k].d;
        }
        if (dir < 3) {
#pragma unroll
          for (int k = 0; k < 9; ++k) {
            w.dRb[dir][k] = R[k].d;
            w.dSz[dir][k] = S[k].d;
          }
        }
      }
    } else if (it >= J0) {
      const int i = it - J0 + 1;
      const double q = x[5 + i];
      const double* ax = m.axis[i];
      double s, c;
      sincos(q, &s, &c);
      const double t = 1.0 - c;
      const double Rq[9] = {t * ax[0] * ax[0] + c,         t * ax[0] * ax[1] - s * ax[2], t * ax[0] * ax[2] + s * ax[1],
                            t * ax[0] * ax[1] + s * ax[2], t * ax[1] * ax[1] + c,         t * ax[1] * ax[2] - s * ax[0],
                            t * ax[0] * ax[2] - s * ax[1], t * ax[1] * ax[2] + s * ax[0], t * ax[2] * ax[2] + c};
      mm3(m.jR[i], Rq, w.Rj[i]);
    }
  }
}

// ---- phase 1b: body-parallel kinematics and body dynamics in ONE pass.  In a common (pelvis) coordinate frame every per-body quantity
// is a sum / product over the ancestor path of the body (<= 7 joints):
//   R_i = prod Rj[k],  p_i = sum R_parent(k) jp_k,  S_k = [p_k x w_k ; w_k],  v_i = v0 + sum S_k qd_k,
//   psd_k = v_parent(k) x S_k,  a_i = a0 + sum (S_k qdd_k + psd_k qd_k),  psdd_k = a_parent(k) x S_k + v_parent(k) x psd_k
// so one thread per body walks its own path, recomputing its ancestors' cheap quantities instead of waiting for them behind barriers
// (this replaces five barrier-separated sweeps), and finishes with the body's spatial inertia and force in the same item.
template <bool DERIV = true>
HD void dynPhaseBodies(Par P, const WbDeviceModel& m, const double* x, const double* u, DynWs& w) {
  for (int i = P.tid; i < NB; i += P.nt) {
    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    V3 p = mk(0, 0, 0);
    V6 v = ld6(w.v0), a = ld6(w.a0);
    V6 S{mk(0, 0, 0), mk(0, 0, 0)}, psd = S, psdd = S;
    const int len = m.pathLen[i];
    for (int t = 0; t < len; ++t) {
      const int k = m.path[i][t];
      p = p + mv(R, ld3(m.jp[k]));
      double T[9];
      mm3(R, w.Rj[k], T);
#pragma unroll
      for (int e = 0; e < 9; ++e) R[e] = T[e];
      const V3 om = mv(R, ld3(m.axis[k]));
      S = V6{cross(p, om), om};
      const double qd = x[NV + 5 + k], qdd = u[12 + k - 1];
      psd = mcross(v, S);                      // v, a still hold the parent's values here
      psdd = mcross(a, S) + mcross(v, psd);
      v = v + qd * S;
      a = a + qdd * S + qd * psd;
    }
#pragma unroll
    for (int e = 0; e < 9; ++e) w.R[i][e] = R[e];
    st3(w.p[i], p);
    st6(w.S[i], S);
    st6(w.v[i], v);
    st6(w.a[i], a);
    if (DERIV) {
      st6(w.psd[i], psd);
      st6(w.psdd[i], psdd);
    }
    // spatial inertia about the pelvis origin and the body force f = I a + v x* I v
    const V3 c = p + mv(R, ld3(m.com[i]));
    double T[9], Ir[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int k = 0; k < 3; ++k) T[3 * r + k] = R[3 * r] * m.Icom[i][k] + R[3 * r + 1] * m.Icom[i][3 + k] + R[3 * r + 2] * m.Icom[i][6 + k];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int k = 0; k < 3; ++k) Ir[3 * r + k] = T[3 * r] * R[3 * k] + T[3 * r + 1] * R[3 * k + 1] + T[3 * r + 2] * R[3 * k + 2];
    const double ms = m.mass[i];
    const double C[9] = {0, -c.z, c.y, c.z, 0, -c.x, -c.y, c.x, 0};
    double* I = DERIV ? w.I[i] : &w.I[0][0] + 9 * i;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double cc = C[3 * r] * C[k] + C[3 * r + 1] * C[3 + k] + C[3 * r + 2] * C[6 + k];
        if (DERIV) {
          I[6 * r + k] = (r == k) ? ms : 0.0;
          I[6 * r + 3 + k] = -ms * C[3 * r + k];
          I[6 * (3 + r) + k] = ms * C[3 * r + k];
          I[6 * (3 + r) + 3 + k] = Ir[3 * r + k] - ms * cc;
        } else {
          I[3 * r + k] = Ir[3 * r + k] - ms * cc;
        }
      }
    // f = I a + v x* (I v), with I = [m 1, -m c^ ; m c^, Ic] applied in closed form
    const V3 hl = ms * (v.l - cross(c, v.a)), ha = ms * cross(c, v.l) + mvIr(Ir, ms, c, v.a);
    const V3 fl = ms * (a.l - cross(c, a.a)), fa = ms * cross(c, a.l) + mvIr(Ir, ms, c, a.a);
    st6(w.f[i], V6{fl, fa} + fcross(v, V6{hl, ha}));
  }
}

// ---- phase 3: B_i = 1/2 [ (v x*) I - I (v x) + (I v) xbar* ]  (items = body x column) ------------------------------------------
HD void dynPhaseBmat(Par P, DynWs& w) {
  for (int it = P.tid; it < NB * 6; it += P.nt) {
    const int i = it / 6, j = it % 6;
    const double* I = w.I[i];
    const V6 v = ld6(w.v[i]);
    const V6 h = m6v(I, v);
    const V6 e = basis6(j);
    V6 Ie{mk(I[j], I[6 + j], I[12 + j]), mk(I[18 + j], I[24 + j], I[30 + j])};
    const V6 col = 0.5 * (fcross(v, Ie) - m6v(I, mcross(v, e)) + fcross(e, h));
    double o[6];
    st6(o, col);
    for (int r = 0; r < 6; ++r) w.Bm[i][6 * r + j] = o[r];
  }
}

// ---- phase 4: composites = sums over subtrees (common coordinates), folded in place: one item per matrix entry, children added to
// parents in decreasing body order (bodies are numbered parent-before-child).  After this phase I, Bm, f hold IC, BC, fC.
// The value-only path needs the root composite only.
template <bool DERIV>
HD void dynPhaseComposite(Par P, const WbDeviceModel& m, DynWs& w) {
  if (DERIV) {
    for (int e = P.tid; e < 78; e += P.nt) {
      if (e < 36) {
        for (int i = NB - 1; i >= 1; --i) w.I[m.parent[i]][e] += w.I[i][e];
      } else if (e < 72) {
        const int k = e - 36;
        for (int i = NB - 1; i >= 1; --i) w.Bm[m.parent[i]][k] += w.Bm[i][k];
      } else {
        const int k = e - 72;
        for (int i = NB - 1; i >= 1; --i) w.f[m.parent[i]][k] += w.f[i][k];
      }
    }
  } else {
    for (int e = P.tid; e < 15; e += P.nt) {   // rotational block of the root composite inertia (packed [NB][9]) and the total force
      double s = 0.0;
      if (e < 9) {
        double* const Iv = &w.I[0][0];
        for (int i = 0; i < NB; ++i) s += Iv[9 * i + e];
        Iv[e] = s;
      } else {
        for (int i = 0; i < NB; ++i) s += w.f[i][e - 9];
        w.f[0][e - 9] = s;
      }
    }
  }
}

// ---- phase 5: net wrench, base acceleration (single item) ----------------------------------------------------------------------------
template <bool DERIV = true>
HD void dynPhaseFinal(Par P, const WbDeviceModel& m, const double* u, DynWs& w) {
  if (P.tid != 0) return;
  V6 E{mk(0, 0, 0), mk(0, 0, 0)};
  for (int c = 0; c < 2; ++c) {
    const int b = m.frameBody[3 * c];
    const V3 pc = ld3(w.p[b]) + mv(w.R[b], ld3(m.frameP[3 * c]));
    const V3 lf = mtv(w.Rb, ld3(u + 6 * c)), lm = mtv(w.Rb, ld3(u + 6 * c + 3));
    st3(w.pc[c], pc);
    st3(w.lf[c], lf);
    st3(w.lm[c], lm);
    E.l = E.l + lf;
    E.a = E.a + cross(pc, lf) + lm;
  }
  const V6 N = E - ld6(w.f[0]);
  st6(w.N, N);
  double Ic[9];
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 3; ++k) Ic[3 * r + k] = DERIV ? w.I[0][6 * (3 + r) + 3 + k] : (&w.I[0][0])[3 * r + k];
  inv3(Ic, w.IcInv);
  const V3 y = mv(w.IcInv, N.a);
  st3(w.y, y);
  st3(w.qddb, (1.0 / m.mtot) * mv(w.Rb, N.l));
  st3(w.qddb + 3, mv(w.SzInv, y));
}

// xdot (58) from the workspace
HD void dynWriteFlow(Par P, const double* x, const double* u, const DynWs& w, double* xdot) {
  for (int i = P.tid; i < NX; i += P.nt) xdot[i] = (i < NV) ? x[NV + i] : (i < NV + 6 ? w.qddb[i - NV] : u[12 + i - NV - 6]);
}

// d(qdd_b)/dz column for direction d given the tangents of the net wrench / composite inertia / base rotations
HD void baseAccTangent(const WbDeviceModel& m, const DynWs& w, V6 dN, const double* dIc, const double* dRb, const double* dSz, double* col) {
  V3 lin = mv(w.Rb, dN.l);
  if (dRb) lin = lin + mv(dRb, ld3(w.N));
  st3(col, (1.0 / m.mtot) * lin);
  V3 rhs = dN.a;
  if (dIc) rhs = rhs - mv(dIc, ld3(w.y));
  V3 dy = mv(w.IcInv, rhs);
  if (dSz) dy = dy - mv(dSz, ld3(w.qddb + 3));
  st3(col + 3, mv(w.SzInv, dy));
}

// ---- phase 6: G = d(qdd_b)/d[x;u]  (6 x 93, column-major with leading dimension 6; one item per column) ------------------------------
// Items are laid out so that the lanes of one warp differentiate the same kind of variable: items 0-31 joint positions, 32-63 joint
// velocities, 64-95 joint accelerations, 96-127 the 24 base / wrench columns (each group padded to 32).
HD void dynPhaseJacobian(Par P, const WbDeviceModel& m, const DynWs& w, double* G) {
  for (int it = P.tid; it < 128; it += P.nt) {
    const int grp = it >> 5, l = it & 31;
    int d;
    if (grp < 3) {
      if (l >= NJ) continue;
      d = (grp == 0 ? 6 : (grp == 1 ? NV + 6 : NX + 12)) + l;
    } else {
      if (l >= 24) continue;
      d = l < 6 ? l : (l < 12 ? NV + (l - 6) : NX + (l - 12));
    }
    double* col = G + 6 * d;
    if (d < 3) {  // base position: no dependence
      for (int k = 0; k < 6; ++k) col[k] = 0.0;
      continue;
    }
    V6 dN{mk(0, 0, 0), mk(0, 0, 0)};
    if (d < 6 || (d >= NV && d < NV + 6)) {
      // base directions: th (d-3), pdot (d-NV), thdot (d-NV-3)
      const int dir = d < 6 ? d - 3 : 3 + (d - NV);
      const V6 dv0 = ld6(w.dv0[dir]), da0 = ld6(w.da0[dir]);
      const V6 v0 = ld6(w.v0);
      const V6 dF = 2.0 * m6v(w.Bm[0], dv0) + m6v(w.I[0], mcross(v0, dv0) + da0);
      dN = V6{-dF.l, -dF.a};
      if (d < 6) {
        const double* dR = w.dRb[dir];
        for (int c = 0; c < 2; ++c) {
          // local force/moment tangents: d(R') f
          const V3 fw = mv(w.Rb, ld3(w.lf[c])), mw = mv(w.Rb, ld3(w.lm[c]));  // back to world values
          const V3 dlf = mtv(dR, fw), dlm = mtv(dR, mw);
          dN.l = dN.l + dlf;
          dN.a = dN.a + cross(ld3(w.pc[c]), dlf) + dlm;
        }
        baseAccTangent(m, w, dN, nullptr, dR, w.dSz[dir], col);
      } else {
        baseAccTangent(m, w, dN, nullptr, nullptr, nullptr, col);
      }
    } else if (d < NV) {  // joint position q_k
      const int k = d - 6 + 1;
      const V6 S = ld6(w.S[k]);
      const V6 dF = fcross(S, ld6(w.f[k])) + m6v(w.I[k], ld6(w.psdd[k])) + 2.0 * m6v(w.Bm[k], ld6(w.psd[k]));
      dN = V6{-dF.l, -dF.a};
      for (int c = 0; c < 2; ++c) {
        const int b = m.frameBody[3 * c];
        if (m.subtree[k] >> b & 1u) {  // contact point moves with joint k
          const V3 dp = cross(S.a, ld3(w.pc[c]) - ld3(w.p[k]));
          dN.a = dN.a + cross(dp, ld3(w.lf[c]));
        }
      }
      // d Ic = rot block of (S x* IC - IC S x)
      const double* X = w.I[k];
      double dIc[9];
      const double sx[9] = {0, -S.l.z, S.l.y, S.l.z, 0, -S.l.x, -S.l.y, S.l.x, 0};
      const double wx[9] = {0, -S.a.z, S.a.y, S.a.z, 0, -S.a.x, -S.a.y, S.a.x, 0};
      for (int r = 0; r < 3; ++r)
        for (int cc = 0; cc < 3; ++cc) {
          double s = 0.0;
          for (int t = 0; t < 3; ++t) {
            s += sx[3 * r + t] * X[6 * t + 3 + cc];            // [s]x M12
            s += wx[3 * r + t] * X[6 * (3 + t) + 3 + cc];      // [w]x M22
            s -= X[6 * (3 + r) + t] * sx[3 * t + cc];          // M21 [s]x
            s -= X[6 * (3 + r) + 3 + t] * wx[3 * t + cc];      // M22 [w]x
          }
          dIc[3 * r + cc] = s;
        }
      baseAccTangent(m, w, dN, dIc, nullptr, nullptr, col);
    } else if (d < NX) {  // joint velocity qd_k
      const int k = d - NV - 6 + 1;
      const V6 dF = 2.0 * (m6v(w.I[k], ld6(w.psd[k])) + m6v(w.Bm[k], ld6(w.S[k])));
      dN = V6{-dF.l, -dF.a};
      baseAccTangent(m, w, dN, nullptr, nullptr, nullptr, col);
    } else if (d < NX + 12) {  // contact wrench components
      const int c = (d - NX) / 6, j = (d - NX) % 6;
      const V3 e = mk(j % 3 == 0, j % 3 == 1, j % 3 == 2);
      const V3 le = mtv(w.Rb, e);
      if (j < 3) {
        dN.l = le;
        dN.a = cross(ld3(w.pc[c]), le);
      } else {
        dN.a = le;
      }
      baseAccTangent(m, w, dN, nullptr, nullptr, nullptr, col);
    } else {  // joint acceleration qdd_k
      const int k = d - NX - 12 + 1;
      const V6 dF = m6v(w.I[k], ld6(w.S[k]));
      dN = V6{-dF.l, -dF.a};
      baseAccTangent(m, w, dN, nullptr, nullptr, nullptr, col);
    }
  }
}

}  // namespace b200sqp
