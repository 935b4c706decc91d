// K1: linear-quadratic approximation of one shooting node (dynamics, costs, constraints, projection), as phase functions.
//
// Replaces, per node, SqpSolver::setupQuadraticSubproblem's body (lib/ocs2_ros2/ocs2_sqp/ocs2_sqp/src/SqpSolver.cpp:346-431):
//   multiple_shooting::setupIntermediateNode (ocs2_oc/src/multiple_shooting/Transcription.cpp:40-94)
//     rk4SensitivityDiscretization (ocs2_core/src/integration/SensitivityIntegratorImpl.cpp:130-169)
//     approximateCost (ocs2_oc/src/approximate_model/LinearQuadraticApproximator.cpp:181-209) * dt
//     equalityConstraintPtr->getLinearApproximation
//   computePerformanceIndex (PerformanceIndexComputation.cpp:40-58)
//   projectTranscription (Transcription.cpp:96-123): luConstraintProjection + changeOfInputVariables
// with the whole-body terms wired as in humanoid_nmpc/humanoid_wb_mpc/src/WBMpcInterface.cpp:131-199.
//
// Structure exploited (B200 formulation): the continuous Jacobian of xdot = [v; qdd_b(x,u); qdd_j] has only six dense rows, so
// the RK4 sensitivity chain carries a 6 x 93 block per stage instead of 58 x 58 / 58 x 35 products.
#pragma once
#include "dense_par.cuh"
#include "wb_feet.cuh"

namespace b200sqp {

// Gauss-Newton / penalty rows of one node, by structure:
constexpr int JS_MAX = 30;     // dense rows: 15 weighted residual rows (quantities 3..17) per swinging foot
constexpr int JU_MAX = 24;     // structured rows: 4 contact-moment rows per stance foot + 16 foot-collision rows (when a foot swings)
constexpr int NUC = 27;        // column support of the structured rows: th (3), the 12 leg joint positions, the 12 contact wrench entries
HD int ucolToZ(int i) { return i < 15 ? 3 + i : NX + (i - 15); }
HD int zToUcol(int d) { return (d >= 3 && d < 18) ? d - 3 : ((d >= NX && d < NX + 12) ? 15 + d - NX : -1); }

struct NodeIn {  // per-node inputs (see b200sqp_upload_instances)
  const double *x, *u, *xnext, *xref;
  double dt;
  int event;          // AnnotatedTime::Event of this node
  int contact[2];
  double swing[2][3], impact[2], armPhase;
  int terminal;
};

// Intermediate record written by K1a (node physics) and consumed by K1b (dense algebra), one per intermediate node, in HBM/L2.
// Offsets in doubles.
struct Mid {
  static constexpr int AB12 = 0;                         // 12 x 93 (ld 12): rows 0-5 base-position rows of [A|B], rows 6-11 base-velocity rows
  static constexpr int B_ = AB12 + 12 * NZ;              // defect b (58)
  static constexpr int CD = B_ + NX;                     // 14 x 93 (ld 14)
  static constexpr int E = CD + NC_MAX * NZ;             // 14
  static constexpr int HDG = E + NC_MAX;                 // Hessian diagonal (93), not yet multiplied by dt
  static constexpr int GQ = HDG + NZ;                    // cost gradient (93)
  static constexpr int FRIC = GQ + NZ;                   // 2 x (3 x 3) friction-cone Hessian blocks
  static constexpr int JS = FRIC + 18;                   // dense weighted rows x 93, leading dimension = their count (15 or 30)
  static constexpr int JU = JS + JS_MAX * NZ;            // structured weighted rows (ld JU_MAX) x NUC compact columns
  static constexpr int META = JU + JU_MAX * NUC;         // nc, swing rows, structured rows, dt, dt*cost
  static constexpr int SIZE = META + 6;
};

struct NodeOut {  // global-memory destinations of one node
  double *A, *Bt, *b, *Q, *St, *Rt, *q, *rt;  // projected stage record in the QP layout (nx = 58, nu_max = 23)
  double *Pu, *Px, *u0;                       // projection u = Pu ut + Px dx + u0
  int* nut;
  double* perf;                               // [dt*cost, dt*|b|^2, dt*|e|^2, projected cost offset]
  double* raw;                                // optional raw (pre-projection) block dump, oracle layout; may be null
};

// per-row scalars of the structured rows
struct RowWs {
  double *wsq, *coef, *val;   // sqrt of the penalty curvature, gradient coefficient of the weighted row, penalty value  (JU_MAX each)
  double* aux;                // collision rows: unit separation direction [JU_MAX][3]
  double* Rf;                 // stance feet: world rotation of the foot frame [2][9]
};
HD constexpr size_t rowWsDoubles() { return 3 * JU_MAX + 3 * JU_MAX + 18; }
HD void rowWsMap(double* base, RowWs& r) {
  r.wsq = base;
  r.coef = r.wsq + JU_MAX;
  r.val = r.coef + JU_MAX;
  r.aux = r.val + JU_MAX;
  r.Rf = r.aux + 3 * JU_MAX;
}

// shared-memory map of K1a (node physics)
struct LqWs {
  DynWs* dyn;
  double* G;       // 4 x (6 x 93)
  double *JFl, *FP, *DFP, *tmpG;
  double* fs;      // 4 x 58 stage flows, b (58), stage point (58)
  double* FV;      // 2 x 18 foot values
  double *gq, *gfoot, *ev, *pv, *sc;   // gfoot: 2 x 93 swing-foot cost gradients ; sc[4..9]: friction gradient
  double *valS, *JU;                   // swing-row values (2 x 18) ; structured rows (aliases dyn->Bm, free between stage 0 and stage 1)
  RowWs rw;
};
// Three nodes per SM was the limit of a 66.7 KB workspace; K1a is bound by the number of resident nodes (see K3), so the map overlays what is
// never alive together: the local foot tangents JFl and the swing-foot gradients gfoot live only during the stage-0 phases and sit on the
// Jacobian blocks of the stage points 1..3, which are written afterwards; the chain scratch tmpG sits on the body inertias of the dynamics
// workspace, dead once the last stage Jacobian exists.  54.9 KB: four nodes per SM (with the model read through L1, not copied per CTA).
HD size_t lqWsDoubles() {
  const size_t dynD = (sizeof(DynWs) + 7) / 8;
  return dynD + 4 * 6 * NZ + 3 * NFRAMES + NFRAMES * 15 * 3 + 6 * NX + 2 * FQ + NZ + NC_MAX + 96 + 16 + 2 * FQ + rowWsDoubles() + 9;
}
HD void lqWsMap(double* base, LqWs& s) {
  const size_t dynD = (sizeof(DynWs) + 7) / 8;
  s.dyn = reinterpret_cast<DynWs*>(base);
  s.G = base + dynD;
  s.JFl = s.G + 6 * NZ;                 // over G[1..3] (stage 0 only)
  s.gfoot = s.JFl + 2 * FLOC * FQ;      // likewise
  static_assert(2 * FLOC * FQ + 2 * NZ <= 3 * 6 * NZ, "JFl and gfoot must fit the stage blocks 1..3");
  s.tmpG = &s.dyn->I[0][0];             // chain phases only: the dynamics workspace is dead by then
  static_assert(6 * NZ <= NB * 36, "tmpG must fit the inertia alias");
  s.FP = s.G + 4 * 6 * NZ;
  s.DFP = s.FP + 3 * NFRAMES;
  s.fs = s.DFP + NFRAMES * 15 * 3;
  s.FV = s.fs + 6 * NX;
  s.gq = s.FV + 2 * FQ;
  s.ev = s.gq + NZ;
  s.pv = s.ev + NC_MAX;
  s.sc = s.pv + 96;
  s.valS = s.sc + 16;
  rowWsMap(s.valS + 2 * FQ, s.rw);
  s.JU = &s.dyn->Bm[0][0];
  static_assert(JU_MAX * NUC <= NB * 36, "structured rows must fit the Bm alias");
}

// Leading dimension of the LU work matrix (nc <= 14 rows x 35 columns, lane = column): odd, so that the 32 lanes walking a row hit distinct
// shared-memory banks (ld 14 put lanes j and j + 8 on the same bank).
constexpr int LU_LD = 15;

// shared-memory map of K1b (projection + change of variables): <= 113 KB so that two CTAs share an SM.
// The change of variables works in the pivoted variable order of the LU (u = [pivot vars (nc) ; free vars (nut)]), where
//   Px = [X ; 0],  u0 = [x0 ; 0],  Pu = [K ; I]      (X | x0 = Xt, K = Kt)
// so every contraction over the 35 inputs shrinks to one over the nc <= 14 pivot variables.
struct PjWs {
  double *Q, *S, *R;             // 58 x 58, 35 x 58 (ld 35), 35 x 35
  double *gq, *bvec;             // 93, 58
  double *Xt, *Kt;               // nc x 59 (ld 14): [X | x0] ; nc x nut (ld 14)
  double* AB12;                  // 12 x 93 (ld 12)
  double* JU;                    // structured rows, JU_MAX x NUC (ld JU_MAX)
  double* scratch;
  int* iw;                       // rowOf[16] colOf[36] posOf[36]
  // phase-1 views (projection)
  double *CD, *ev, *LU;
  double* JS1;                   // dense rows of ONE swinging foot, 15 x 93 (ld 15), staged next to CD | e | LU (CUDA path)
  // dynamics views
  double *B1, *D12;              // 12 x nc (ld 12) ; 12 x (59 + nut) (ld 12)
  // Hessian view
  double* JRc;                   // swing rows, JS_MAX x 93 (ld JS_MAX)
  // cost change-of-variables views
  double *T11, *T12, *R11, *R21, *W, *V, *rr;   // nc x 58 (ld 14), nut x 58 (ld 23), nc x nc (ld 14), nut x nc (ld 23), nc x nut (ld 14), nut x nut (ld 23), 35 + 35
};
constexpr int PJ_SCRATCH = 3600;   // max over the phases: 1841 + 1395 (projection + staged swing rows), 1152 (dynamics), 2790 (swing rows), 3585 (cost change of variables)
HD size_t pjWsDoubles() {
  return NX * NX + NU * NX + NU * NU + 1 + NZ + 1 + NX + NC_MAX * (NX + 1) + NC_MAX * NUT_MAX + 12 * NZ + JU_MAX * NUC + PJ_SCRATCH + 48;
}
HD void pjWsMap(double* base, PjWs& s) {
  s.Q = base;
  s.S = s.Q + NX * NX;
  s.R = s.S + NU * NX;
  s.gq = s.R + NU * NU + 1;
  s.bvec = s.gq + NZ + 1;
  s.Xt = s.bvec + NX;
  s.Kt = s.Xt + NC_MAX * (NX + 1);
  s.AB12 = s.Kt + NC_MAX * NUT_MAX;
  s.JU = s.AB12 + 12 * NZ;
  s.scratch = s.JU + JU_MAX * NUC;
  s.iw = reinterpret_cast<int*>(s.scratch + PJ_SCRATCH);
  // phase 1 (projection): CD | e | LU | triangular-solve workspace
  s.CD = s.scratch;                          // 14 x 93 = 1302
  s.ev = s.CD + NC_MAX * NZ;                 // 14
  s.LU = s.ev + NC_MAX;                      // 14 x 35 with ld LU_LD = 525
  s.JS1 = s.LU + LU_LD * NU;                 // 15 x 93 = 1395   (ends at 3236)
  // dynamics change of variables
  s.B1 = s.scratch;                          // 12 x 14 = 168
  s.D12 = s.B1 + 12 * NC_MAX;                // 12 x 82 = 984
  // Hessian
  s.JRc = s.scratch;                         // 30 x 93 = 2790
  // cost change of variables
  s.T11 = s.scratch;                         // 14 x 58 = 812
  s.T12 = s.T11 + NC_MAX * NX;               // 23 x 58 = 1334
  s.R11 = s.T12 + NUT_MAX * NX;              // 196
  s.R21 = s.R11 + NC_MAX * NC_MAX;           // 23 x 14 = 322
  s.W = s.R21 + NUT_MAX * NC_MAX;            // 14 x 23 = 322
  s.V = s.W + NC_MAX * NUT_MAX;              // 23 x 23 = 529
  s.rr = s.V + NUT_MAX * NUT_MAX;            // 70   (ends at 3585)
}

// shared-memory map of K1b part 2 when Q accumulates in its final place in HBM/L2 (wb_node_b2.inc): no Q, no dynamics operands -- 70 KB,
// three nodes per SM instead of two
HD size_t pjCostWsDoubles() { return NU * NX + NU * NU + 1 + NZ + 1 + NC_MAX * (NX + 1) + NC_MAX * NUT_MAX + JU_MAX * NUC + PJ_SCRATCH + 48; }
HD void pjCostWsMap(double* base, double* Qglobal, PjWs& s) {
  pjWsMap(base, s);   // sets the scratch views relative to s.scratch below
  s.Q = Qglobal;
  s.S = base;
  s.R = s.S + NU * NX;
  s.gq = s.R + NU * NU + 1;
  s.bvec = nullptr;
  s.AB12 = nullptr;
  s.Xt = s.gq + NZ + 1;
  s.Kt = s.Xt + NC_MAX * (NX + 1);
  s.JU = s.Kt + NC_MAX * NUT_MAX;
  double* const scratch = s.JU + JU_MAX * NUC;
  const ptrdiff_t shift = scratch - s.scratch;
  s.scratch = scratch;
  s.iw = reinterpret_cast<int*>(s.scratch + PJ_SCRATCH);
  s.CD += shift; s.ev += shift; s.LU += shift; s.JS1 += shift; s.B1 += shift; s.D12 += shift; s.JRc += shift;
  s.T11 += shift; s.T12 += shift; s.R11 += shift; s.R21 += shift; s.W += shift; s.V += shift; s.rr += shift;
}

// shared-memory map of K1b part 1 (projection + dynamics, wb_node_b1.inc): the same PjWs views on a 34 KB workspace without the Hessian blocks
HD size_t pjDynWsDoubles() { return NC_MAX * (NX + 1) + NC_MAX * NUT_MAX + 12 * NZ + NX + (NC_MAX * NZ + NC_MAX + LU_LD * NU) + 48; }
HD void pjDynWsMap(double* base, PjWs& s) {
  s.Q = s.S = s.R = s.gq = s.JU = s.JS1 = s.JRc = nullptr;
  s.T11 = s.T12 = s.R11 = s.R21 = s.W = s.V = s.rr = nullptr;
  s.Xt = base;
  s.Kt = s.Xt + NC_MAX * (NX + 1);
  s.AB12 = s.Kt + NC_MAX * NUT_MAX;
  s.bvec = s.AB12 + 12 * NZ;
  s.scratch = s.bvec + NX;
  s.CD = s.scratch;                          // 14 x 93
  s.ev = s.CD + NC_MAX * NZ;                 // 14
  s.LU = s.ev + NC_MAX;                      // 14 x 35 with ld LU_LD
  s.B1 = s.scratch;                          // 12 x 14, over the dead CD
  s.D12 = s.B1 + 12 * NC_MAX;                // 12 x 82
  s.iw = reinterpret_cast<int*>(s.LU + LU_LD * NU);
}

HD void penRelaxed(double mu, double delta, double h, double& v, double& d1, double& d2) {
  // RelaxedBarrierPenalty (ocs2_core/src/penalties/penalties/RelaxedBarrierPenalty.cpp:37-66)
  if (h > delta) {
    v = -mu * log(h);
    d1 = -mu / h;
    d2 = mu / (h * h);
  } else {
    const double dh = (h - 2.0 * delta) / delta;
    v = mu * (-log(delta) + 0.5 * dh * dh - 0.5);
    d1 = mu * ((h - 2.0 * delta) / (delta * delta));
    d2 = mu / (delta * delta);
  }
}
HD void penPwPoly(double mu, double delta, double h, double& v, double& d1, double& d2) {
  // PieceWisePolynomialBarrierPenalty (.../PieceWisePolynomialBarrierPenalty.cpp:37-75)
  if (h <= 0) {
    v = mu * (0.5 * h * h - delta * h / 2 + delta * delta / 6);
    d1 = mu * (h - delta / 2);
    d2 = mu;
  } else if (h < delta) {
    v = mu * (-h * h * h / (6 * delta) + 0.5 * h * h - delta * h / 2 + delta * delta / 6);
    d1 = mu * (-h * h / (2 * delta) + h - delta / 2);
    d2 = mu * (-h / delta + 1);
  } else {
    v = d1 = d2 = 0.0;
  }
}

// ---- RK4 sensitivity chain on the structured Jacobian ----------------------------------------------------------------------------------
// Gr_s = Gq_s Xq_s + Gv_s Xv_s + Gu_s with  Xq_1 = e_q, Xv_1 = e_v,  Xq_s = e_q + c_s Xv_{s-1},  Xv_s = e_v + c_s [Gr_{s-1}; e_aj].
// In place on G: after the call G[s] holds Gr_s (6 x 93, ld 6).  One item per (stage processed sequentially by the caller, column).
HD double xvEntry(const double* GrPrev, double c, int row, int d) {  // Xv_s(row, d), rows = generalized velocities (29)
  double e = (d == NV + row) ? 1.0 : 0.0;
  if (c != 0.0) e += c * (row < 6 ? GrPrev[row + 6 * d] : ((d == NX + 12 + row - 6) ? 1.0 : 0.0));
  return e;
}
HD void rkPhaseChainStage(Par P, int s, double dt, double* G, double* tmp /*6 x 93*/) {
  // stage point x_s = x + c_s k_{s-1}; c = c_s, cp = c_{s-1}.  With the sparse structure of Xq, Xv each entry needs 12 products:
  //   Gq Xq = Gq[:,d][d<29] + c ( Gq[:,d-29][29<=d<58] + cp ( Gq[:,0:6] Gr_{s-2}[:,d] + Gq[:,6+d-70][d>=70] ) )
  //   Gv Xv = Gv[:,d-29][29<=d<58] + c ( Gv[:,0:6] Gr_{s-1}[:,d] + Gv[:,6+d-70][d>=70] )
  const double c = (s == 3) ? dt : 0.5 * dt;
  const double cp = (s == 1) ? 0.0 : 0.5 * dt;
  const double* Gs = G + s * 6 * NZ;
  const double* GrP = G + (s - 1) * 6 * NZ;
  const double* GrPP = (s >= 2) ? G + (s - 2) * 6 * NZ : nullptr;
  for (int it = P.tid; it < 6 * NZ; it += P.nt) {
    const int r = it % 6, d = it / 6;
    const double* Gq = Gs + r;                 // Gq[r][k] = Gq[6 * k]
    const double* Gv = Gs + r + 6 * NV;
    double acc = (d >= NX) ? Gs[r + 6 * d] : 0.0;
    if (d < NV) acc += Gq[6 * d];
    double inner_q = 0.0, inner_v = 0.0;
    if (d >= NV && d < NX) {
      inner_q = Gq[6 * (d - NV)];
      acc += Gv[6 * (d - NV)];
    }
    if (d >= NX + 12) {
      inner_v = Gv[6 * (6 + d - NX - 12)];
      if (cp != 0.0) inner_q = fma(cp, Gq[6 * (6 + d - NX - 12)], inner_q);
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) inner_v = fma(Gv[6 * k], GrP[k + 6 * d], inner_v);
    if (cp != 0.0) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) t = fma(Gq[6 * k], GrPP[k + 6 * d], t);
      inner_q = fma(cp, t, inner_q);
    }
    tmp[it] = fma(c, inner_q + inner_v, acc);
  }
}

// The 12 dense rows of [A | B] (A = I + sum_s w_s dk_s/dx, B = sum_s w_s dk_s/du): base position rows 0-5 and base velocity rows 29-34;
// every other row is a fixed pattern (q_j+ = q_j + dt v_j + dt^2/2 a_j, v_j+ = v_j + dt a_j).  b = x + sum w_s k_s - xnext.
HD void rkPhaseAssemble12(Par P, double dt, const double* x, const double* xnext, const double* G, const double* fs, double* AB12, double* b) {
  const double w[4] = {dt / 6.0, dt / 3.0, dt / 3.0, dt / 6.0};
  const double cs[4] = {0.0, 0.5 * dt, 0.5 * dt, dt};
  for (int it = P.tid; it < 12 * NZ; it += P.nt) {
    const int rr = it % 12, d = it / 12;
    const int r = rr < 6 ? rr : NV + (rr - 6);
    double acc = (r == d) ? 1.0 : 0.0;
    if (rr < 6) {
      for (int s = 0; s < 4; ++s) acc = fma(w[s], xvEntry(s > 0 ? G + (s - 1) * 6 * NZ : nullptr, cs[s], r, d), acc);
    } else {
      for (int s = 0; s < 4; ++s) acc = fma(w[s], G[s * 6 * NZ + (rr - 6) + 6 * d], acc);
    }
    AB12[it] = acc;
  }
  for (int r = P.tid; r < NX; r += P.nt) {
    double acc = x[r];
    for (int s = 0; s < 4; ++s) acc = fma(w[s], fs[s * NX + r], acc);
    b[r] = acc - xnext[r];
  }
}
// dense entry (r, d) of [A | B] from the 12 stored rows and the fixed pattern
HD double abEntry(const double* AB12, double dt, int r, int d) {
  if (r < 6) return AB12[r + 12 * d];
  if (r >= NV && r < NV + 6) return AB12[6 + (r - NV) + 12 * d];
  if (r < NV) {  // joint position row
    const int j = r - 6;
    return (d == r ? 1.0 : 0.0) + (d == NV + r ? dt : 0.0) + (d == NX + 12 + j ? 0.5 * dt * dt : 0.0);
  }
  const int j = r - NV - 6;  // joint velocity row
  return (d == r ? 1.0 : 0.0) + (d == NX + 12 + j ? dt : 0.0);
}

// ---- equality constraints C dx + D du + e (collection order of WBMpcInterface.cpp:172-181) ----------------------------------------------
HD int constraintRowCount(const NodeIn& n) { return (n.contact[0] ? 6 : 7) + (n.contact[1] ? 6 : 7); }
HD void conPhaseAssemble(Par P, const WbDeviceModel& m, const NodeIn& n, const double* JF, const double* FV, double* CD, double* ev,
                         bool valuesOnly = false) {
  const int nc = constraintRowCount(n);
  for (int it = P.tid; it < nc * (valuesOnly ? 1 : NZ + 1); it += P.nt) {
    const int r = it % nc, d = valuesOnly ? NZ : it / nc;  // d == NZ -> constant term
    int c = 0, lr = r;
    const int n0 = n.contact[0] ? 6 : 7;
    if (r >= n0) {
      c = 1;
      lr = r - n0;
    }
    const double* J = JF + static_cast<size_t>(c * NZ + (d < NZ ? d : 0)) * FQ;
    const double* F = FV + FQ * c;
    double val;
    if (n.contact[c]) {
      // ZeroAccelerationConstraintCppAd: Ax [p; oriErr] + Av twist + Aa acc   (EndEffectorDynamicsAccelerationsConstraint.cpp:84-146)
      const double Av = lr < 2 ? m.gLinVelXY : (lr == 2 ? m.gLinVelZ : m.gAngVel);
      const double Aa = lr < 2 ? m.gLinAccXY : (lr == 2 ? m.gLinAccZ : m.gAngAcc);
      const double Ax = lr == 2 ? m.gPosZ : (lr >= 3 ? m.gOri : 0.0);
      const double* src = d < NZ ? J : F;
      val = Ax * src[lr] + Av * src[6 + lr] + Aa * src[12 + lr];
    } else if (lr < 6) {
      // ZeroWrenchConstraint
      val = d < NZ ? ((d == NX + 6 * c + lr) ? 1.0 : 0.0) : n.u[6 * c + lr];
    } else {
      // SwingLegVerticalConstraintCppAd (WBMpcPreComputation.cpp:92-103)
      const double* src = d < NZ ? J : F;
      val = m.gPosZ * src[2] + m.gLinVelZ * src[8] + m.gLinAccZ * src[14];
      if (d == NZ) val += -m.gLinVelZ * n.swing[c][1] - m.gLinAccZ * n.swing[c][2] - m.gPosZ * n.swing[c][0];
    }
    if (d < NZ) CD[r + NC_MAX * d] = val;
    else ev[r] = val;
  }
}

// ---- costs ------------------------------------------------------------------------------------------------------------------------------------
// phase A: zero H, diagonal terms, gradient of the quadratic tracking cost, joint limits, friction cone (items over 93 + a few)
template <bool DERIV = true>
HD void costPhaseDiagonal(Par P, const WbDeviceModel& m, const NodeIn& n, double* hdiag, double* gq, double* pv /*partial values, 96*/) {
  // friction-cone penalty derivative (needed for the Hessian shift on every diagonal entry)
  double fricD1[2] = {0.0, 0.0};
  for (int c = 0; c < 2; ++c)
    if (n.contact[c]) {
      const double* F = n.u + 6 * c;
      const double h = m.fricCoeff * F[2] - sqrt(F[0] * F[0] + F[1] * F[1] + m.fricReg);
      double v, d2;
      penRelaxed(m.fricMu, m.fricDelta, h, v, fricD1[c], d2);
    }
  const double shift = -(fricD1[0] + fricD1[1]) * m.fricShift;
  for (int i = P.tid; i < NZ; i += P.nt) {
    double val = 0.0, g = 0.0, hd = shift;
    if (i < NX) {
      // StateInputQuadraticCost about xNominal (arm-swing reference evaluated at the current yaw, treated as constant)
      double xn = n.xref[i];
      if (i >= 6 && i < NV) {
        const int j = i - 6;
        const double yaw = n.x[3];
        const double gc = n.armPhase * (cos(yaw) * n.xref[NV] + sin(yaw) * n.xref[NV + 1]);
        if (j == m.armJoint[0] || j == m.armJoint[2]) xn += -0.15 * gc;
        if (j == m.armJoint[1] || j == m.armJoint[3]) xn += 0.15 * gc;
      }
      const double dx = n.x[i] - xn;
      val = 0.5 * m.Qd[i] * dx * dx;
      g = m.Qd[i] * dx;
      hd += m.Qd[i];
      if (i >= 6 && i < NV) {  // JointLimitsSoftConstraint
        const int j = i - 6;
        double vu, d1u, d2u, vl, d1l, d2l;
        penPwPoly(m.jlMu, m.jlDelta, m.qhi[j] - n.x[i], vu, d1u, d2u);
        penPwPoly(m.jlMu, m.jlDelta, n.x[i] - m.qlo[j], vl, d1l, d2l);
        val += vu + vl;
        g += d1l - d1u;
        hd += d2l + d2u;
      }
    } else {
      const int j = i - NX;
      double un = 0.0;
      const int ns = n.contact[0] + n.contact[1];
      if (ns > 0 && (j == 2 || j == 8) && n.contact[j / 6]) un = m.mtot * 9.81 / ns;
      const double du = n.u[j] - un;
      val = 0.5 * m.Rd[j] * du * du;
      g = m.Rd[j] * du;
      hd += m.Rd[j];
    }
    if (DERIV) {
      hdiag[i] = hd;
      gq[i] = g;
    }
    pv[i] = val;
  }
}
// friction cone blocks (items = contacts): value, gradient, 3x3 Hessian block  (FrictionForceConeConstraint.cpp:145-224)
template <bool DERIV = true>
HD void costPhaseFriction(Par P, const WbDeviceModel& m, const NodeIn& n, double* fric /*2 x 9*/, double* fricG /*2 x 3*/, double* pv) {
  for (int c = P.tid; c < 2; c += P.nt) {
    pv[NZ + c] = 0.0;
    if (DERIV) {
      for (int k = 0; k < 9; ++k) fric[9 * c + k] = 0.0;
      for (int k = 0; k < 3; ++k) fricG[3 * c + k] = 0.0;
    }
    if (!n.contact[c]) continue;
    const double* F = n.u + 6 * c;
    const double Ft2 = F[0] * F[0] + F[1] * F[1] + m.fricReg, Ft = sqrt(Ft2), Ft32 = Ft * Ft2;
    const double h = m.fricCoeff * F[2] - Ft;
    double v, d1, d2;
    penRelaxed(m.fricMu, m.fricDelta, h, v, d1, d2);
    const double dh[3] = {-F[0] / Ft, -F[1] / Ft, m.fricCoeff};
    const double ddh[9] = {-(F[1] * F[1] + m.fricReg) / Ft32, F[0] * F[1] / Ft32, 0, F[0] * F[1] / Ft32, -(F[0] * F[0] + m.fricReg) / Ft32, 0, 0, 0, 0};
    pv[NZ + c] = v;
    if (!DERIV) continue;
    for (int i = 0; i < 3; ++i) {
      fricG[3 * c + i] = d1 * dh[i];
      for (int j = 0; j < 3; ++j) fric[9 * c + 3 * i + j] = d2 * dh[i] * dh[j] + d1 * ddh[3 * i + j];
    }
  }
}

// ---- structured penalty rows -------------------------------------------------------------------------------------------------------------------
HD int momRows(const NodeIn& n, int c) { return n.contact[c] ? 4 : 0; }
HD int collRows(const NodeIn& n) { return (n.contact[0] && n.contact[1]) ? 0 : 16; }
HD int structRows(const NodeIn& n) { return momRows(n, 0) + momRows(n, 1) + collRows(n); }
HD int swingRows(const NodeIn& n) { return (n.contact[0] ? 0 : 15) + (n.contact[1] ? 0 : 15); }
HD void collisionPair(int r, int& a, int& b, bool& knee) {
  // FootCollisionConstraint.cpp:112-137 ; frames: 0 fl, 1 fl_p1, 2 fl_p2, 3 fr, 4 fr_p1, 5 fr_p2, 6 ankle_l, 7 ankle_r, 8 knee_l, 9 knee_r
  const int A[16] = {1, 1, 2, 2, 0, 0, 3, 3, 0, 8, 0, 1, 2, 3, 4, 5};
  const int B[16] = {4, 5, 4, 5, 4, 5, 1, 2, 3, 9, 7, 7, 7, 6, 6, 6};
  a = A[r];
  b = B[r];
  knee = (r == 9);
}
// per-row scalars, one item per structured row (row order: moment rows of foot 0, of foot 1, collision rows)
HD void costPhaseRowScalars(Par P, const WbDeviceModel& m, const NodeIn& n, const DynWs& w, const double* FP, RowWs rw) {
  const int nm0 = momRows(n, 0), nm = nm0 + momRows(n, 1), nr = nm + collRows(n);
  for (int it = P.tid; it < nr; it += P.nt) {
    double wsq = 0.0, coef = 0.0, v = 0.0, d1, d2;
    if (it < nm) {
      // ContactMomentXYConstraintCppAd (ContactMomentXYConstraintCppAd.cpp:86-104) under a relaxed barrier:  h_r = sm * lm[ax] + bound * lf.z
      const int c = (it < nm0) ? 0 : 1, r = it - (c ? nm0 : 0);
      double Rf[9];
      mm3(w.Rb, w.R[m.frameBody[3 * c]], Rf);
      const V3 lf = mtv(Rf, ld3(n.u + 6 * c)), lm = mtv(Rf, ld3(n.u + 6 * c + 3));
      const double sm = (r == 0 || r == 3) ? 1.0 : -1.0;
      const double bound = (r == 0) ? -m.rect[2] : (r == 1 ? m.rect[3] : (r == 2 ? -m.rect[0] : m.rect[1]));
      const double h = sm * (r < 2 ? lm.x : lm.y) + bound * lf.z;
      penRelaxed(m.momMu, m.momDelta, h, v, d1, d2);
      wsq = sqrt(d2);
      coef = d1 / wsq;
      if (r == 0)
        for (int k = 0; k < 9; ++k) rw.Rf[9 * c + k] = Rf[k];
    } else {
      // FootCollisionConstraint under the piecewise-polynomial barrier
      int a, b;
      bool knee;
      collisionPair(it - nm, a, b, knee);
      const V3 dv = ld3(FP + 3 * a) - ld3(FP + 3 * b);
      const double dist = sqrt(dot(dv, dv));
      penPwPoly(m.collMu, m.collDelta, dist - 2.0 * (knee ? m.rKnee : m.rFoot), v, d1, d2);
      if (d2 > 0.0) {
        wsq = sqrt(d2);
        coef = d1 / wsq;
      }
      st3(rw.aux + 3 * it, (1.0 / dist) * dv);
    }
    rw.wsq[it] = wsq;
    rw.coef[it] = coef;
    rw.val[it] = v;
  }
}
// weighted row entries over the compact column support, one item per (row, compact column)
HD void costPhaseRowEntries(Par P, const WbDeviceModel& m, const NodeIn& n, const DynWs& w, const double* DFP, RowWs rw, double* JU) {
  const int nm0 = momRows(n, 0), nm = nm0 + momRows(n, 1), nr = nm + collRows(n);
  for (int it = P.tid; it < JU_MAX * NUC; it += P.nt) {
    const int r = it % JU_MAX, i = it / JU_MAX;   // compile-time divisors; rows beyond the active count are skipped
    if (r >= nr) continue;
    double e = 0.0;
    if (r < nm) {
      const int c = (r < nm0) ? 0 : 1, lr = r - (c ? nm0 : 0);
      const int b = m.frameBody[3 * c];
      const double* Rf = rw.Rf + 9 * c;
      const double sm = (lr == 0 || lr == 3) ? 1.0 : -1.0;
      const double bound = (lr == 0) ? -m.rect[2] : (lr == 1 ? m.rect[3] : (lr == 2 ? -m.rect[0] : m.rect[1]));
      if (i < 15) {
        // tangent of the foot rotation: dRf = [omega]x Rf  ->  d(Rf' a) = -Rf'(omega x a)
        V3 om = mk(0, 0, 0);
        bool rotd = false;
        if (i < 3) {
          om = mv(w.Rb, mk(w.Sz[i], w.Sz[3 + i], w.Sz[6 + i]));   // world angular direction of th_i
          rotd = true;
        } else if (m.subtree[i - 2] >> b & 1u) {
          om = mv(w.Rb, ld3(w.S[i - 2] + 3));
          rotd = true;
        }
        if (rotd) {
          const V3 dlf = -mtv(Rf, cross(om, ld3(n.u + 6 * c))), dlm = -mtv(Rf, cross(om, ld3(n.u + 6 * c + 3)));
          e = sm * (lr < 2 ? dlm.x : dlm.y) + bound * dlf.z;
        }
      } else {
        const int j = i - 15 - 6 * c;
        if (j >= 0 && j < 6) {
          const V3 col = mk(Rf[3 * (j % 3)], Rf[3 * (j % 3) + 1], Rf[3 * (j % 3) + 2]);  // Rf' e_j = row j of Rf
          e = (j < 3) ? bound * col.z : sm * (lr < 2 ? col.x : col.y);
        }
      }
    } else if (i < 15) {
      int a, b;
      bool knee;
      collisionPair(r - nm, a, b, knee);
      e = dot(ld3(rw.aux + 3 * r), ld3(DFP + (a * 15 + i) * 3) - ld3(DFP + (b * 15 + i) * 3));
    }
    JU[r + JU_MAX * i] = rw.wsq[r] * e;
  }
}
// values of the swing-foot residual rows: valS[c][r] = 1/2 (sqrtW_r impact_c FV_r)^2   (EndEffectorDynamicsFootCost.cpp:91-124)
HD void costPhaseSwingValues(Par P, const WbDeviceModel& m, const NodeIn& n, const double* FV, double* valS) {
  for (int it = P.tid; it < 2 * FQ; it += P.nt) {
    const int c = it / FQ, r = it % FQ;
    const double res = (n.contact[c] || r < 3) ? 0.0 : m.footSqrtW[r] * n.impact[c] * FV[it];
    valS[it] = 0.5 * res * res;
  }
}

// Fused foot phase, one item per (foot c, tangent direction d): column d of the foot's 18 quantities (scatter of the local tangents + the
// chain through the base acceleration, J_fb G) is formed in registers and consumed at once: the foot's equality-constraint rows go to CD,
// a swinging foot's 18 weighted residual rows to JR, and its share of the cost gradient to gfoot[c][d].  Nothing dense is staged.
HD int footZToLocal(const WbDeviceModel& m, int c, int d) {
  int body, base;
  if (d < 6) return d;                              // pb, th
  if (d < NV) {
    body = d - 6 + 1;                               // q_leg
    base = 6;
  } else if (d < NV + 6) {
    return 12 + (d - NV);                           // pd, thd
  } else if (d < NX) {
    body = d - NV - 6 + 1;                          // qd_leg
    base = 18;
  } else if (d < NX + 12) {
    return -1;                                      // contact wrenches act only through qdd_b
  } else {
    body = d - NX - 12 + 1;                         // qdd_leg
    base = 24;
  }
#pragma unroll
  for (int t = 0; t < LEG_LEN; ++t)
    if (m.legBody[c][t] == body) return base + t;
  return -1;
}
HD void footPhaseColumns(Par P, const WbDeviceModel& m, const NodeIn& n, const double* JFl, const double* G, const double* FV, double* CD,
                         double* JS, double* gfoot) {
  const int n0 = n.contact[0] ? 6 : 7;
  for (int it = P.tid; it < 2 * NZ; it += P.nt) {
    const int c = it / NZ, d = it - c * NZ;
    double col[FQ];
    const int l = footZToLocal(m, c, d);
    const double* src = JFl + (c * FLOC + (l >= 0 ? l : 0)) * FQ;
#pragma unroll
    for (int k = 0; k < FQ; ++k) col[k] = (l >= 0) ? src[k] : 0.0;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const double g = G[j + 6 * d];
      const double* jb = JFl + (c * FLOC + 30 + j) * FQ;
#pragma unroll
      for (int k = 12; k < FQ; ++k) col[k] = fma(jb[k], g, col[k]);
    }
    double* cd = CD + (c ? n0 : 0) + NC_MAX * d;
    double gsum = 0.0;
    if (n.contact[c]) {
      // ZeroAccelerationConstraintCppAd rows (see conPhaseAssemble)
#pragma unroll
      for (int lr = 0; lr < 6; ++lr) {
        const double Av = lr < 2 ? m.gLinVelXY : (lr == 2 ? m.gLinVelZ : m.gAngVel);
        const double Aa = lr < 2 ? m.gLinAccXY : (lr == 2 ? m.gLinAccZ : m.gAngAcc);
        const double Ax = lr == 2 ? m.gPosZ : (lr >= 3 ? m.gOri : 0.0);
        cd[lr] = Ax * col[lr] + Av * col[6 + lr] + Aa * col[12 + lr];
      }
    } else {
#pragma unroll
      for (int lr = 0; lr < 6; ++lr) cd[lr] = (d == NX + 6 * c + lr) ? 1.0 : 0.0;            // ZeroWrenchConstraint
      cd[6] = m.gPosZ * col[2] + m.gLinVelZ * col[8] + m.gLinAccZ * col[14];                 // SwingLegVerticalConstraintCppAd
      double* js = JS + ((c && !n.contact[0]) ? 15 : 0) + swingRows(n) * d;                  // EndEffectorDynamicsFootCost rows 3..17 (ld = row count)
#pragma unroll
      for (int r = 3; r < FQ; ++r) {
        const double sw = m.footSqrtW[r] * n.impact[c];
        const double e = sw * col[r];
        js[r - 3] = e;
        gsum = fma(e, sw * FV[FQ * c + r], gsum);
      }
    }
    gfoot[it] = gsum;
  }
}
// gq += swing-foot gradients + friction gradient + JU' coef ; the structured rows are copied to the record on the way
HD void costPhaseGradient(Par P, const NodeIn& n, const double* JU, const double* coef, const double* gfoot, const double* fricG, double* gq,
                          double* midJU) {
  const int nr = structRows(n);
  for (int i = P.tid; i < NZ; i += P.nt) {
    double acc = gfoot[i] + gfoot[NZ + i];
    if (i >= NX && i < NX + 12 && (i - NX) % 6 < 3) acc += fricG[3 * ((i - NX) / 6) + (i - NX) % 6];
    const int uc = zToUcol(i);
    if (uc >= 0)
      for (int r = 0; r < nr; ++r) acc = fma(JU[r + JU_MAX * uc], coef[r], acc);
    gq[i] += acc;
  }
  for (int it = P.tid; it < JU_MAX * NUC; it += P.nt) midJU[it] = (it % JU_MAX < nr) ? JU[it] : 0.0;
}


// ---- projection: Eigen::FullPivLU semantics (complete pivoting; particular solution with free variables = 0; kernel basis) ----------------
// Whole factorisation by ONE warp: lane l owns columns l and l+32; pivot search = per-lane scan + shuffle arg-max with ties resolved
// towards the smaller column-major index (the first maximum of Eigen's / the oracle's scan).  Host harness: sequential.
HD void luPhaseFactor(Par P, int nc, double* LU, int* rowOf, int* colOf) {
#ifdef __CUDA_ARCH__
  if (P.tid >= 32) return;
  const int lane = P.tid;
  for (int k = 0; k < nc; ++k) {
    double best = -1.0;
    int bidx = 1 << 30;
    for (int j = lane; j < NU; j += 32)
      if (j >= k)
        for (int i = k; i < nc; ++i) {
          const double a = fabs(LU[i + LU_LD * j]);
          if (a > best) {
            best = a;
            bidx = j * NC_MAX + i;
          }
        }
#pragma unroll
    for (int off = 16; off; off >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx, off);
      if (ob > best || (ob == best && oi < bidx)) {
        best = ob;
        bidx = oi;
      }
    }
    const int pj = bidx / NC_MAX, pi = bidx % NC_MAX;
    if (pi != k)
      for (int j = lane; j < NU; j += 32) {
        const double t = LU[k + LU_LD * j];
        LU[k + LU_LD * j] = LU[pi + LU_LD * j];
        LU[pi + LU_LD * j] = t;
      }
    __syncwarp();
    if (pj != k)
      for (int i = lane; i < nc; i += 32) {
        const double t = LU[i + LU_LD * k];
        LU[i + LU_LD * k] = LU[i + LU_LD * pj];
        LU[i + LU_LD * pj] = t;
      }
    if (lane == 0) {
      int t = rowOf[k];
      rowOf[k] = rowOf[pi];
      rowOf[pi] = t;
      t = colOf[k];
      colOf[k] = colOf[pj];
      colOf[pj] = t;
    }
    __syncwarp();
    const double piv = LU[k + LU_LD * k];
    for (int i = k + 1 + lane; i < nc; i += 32) LU[i + LU_LD * k] /= piv;
    __syncwarp();
    for (int j = k + 1 + lane; j < NU; j += 32) {
      const double ukj = LU[k + LU_LD * j];
      for (int i = k + 1; i < nc; ++i) LU[i + LU_LD * j] = fma(-LU[i + LU_LD * k], ukj, LU[i + LU_LD * j]);
    }
    __syncwarp();
  }
#else
  if (P.tid != 0) return;
  for (int k = 0; k < nc; ++k) {
    int pi = k, pj = k;
    double best = -1.0;
    for (int j = k; j < NU; ++j)
      for (int i = k; i < nc; ++i) {
        const double a = fabs(LU[i + LU_LD * j]);
        if (a > best) {
          best = a;
          pi = i;
          pj = j;
        }
      }
    if (pi != k) {
      for (int j = 0; j < NU; ++j) {
        const double t = LU[k + LU_LD * j];
        LU[k + LU_LD * j] = LU[pi + LU_LD * j];
        LU[pi + LU_LD * j] = t;
      }
      const int t = rowOf[k];
      rowOf[k] = rowOf[pi];
      rowOf[pi] = t;
    }
    if (pj != k) {
      for (int i = 0; i < nc; ++i) {
        const double t = LU[i + LU_LD * k];
        LU[i + LU_LD * k] = LU[i + LU_LD * pj];
        LU[i + LU_LD * pj] = t;
      }
      const int t = colOf[k];
      colOf[k] = colOf[pj];
      colOf[pj] = t;
    }
    const double piv = LU[k + LU_LD * k];
    for (int i = k + 1; i < nc; ++i) LU[i + LU_LD * k] /= piv;
    for (int j = k + 1; j < NU; ++j) {
      const double ukj = LU[k + LU_LD * j];
      for (int i = k + 1; i < nc; ++i) LU[i + LU_LD * j] = fma(-LU[i + LU_LD * k], ukj, LU[i + LU_LD * j]);
    }
  }
#endif
}
// Px = -D^+ C (35 x 58), u0 = -D^+ e, Pu = kernel (35 x nut), from the LU factors (Eigen: particular solution with the free variables
// at zero, kernel basis from the pivoted U):
// [X | x0] = -U^-1 L^-1 P [C | e]  (nc x 59)  and  K = -U^-1 U_rk  (nc x nut), one right-hand side per work item: the column lives in
// registers through the unit-lower forward and the upper backward substitution (replaces explicit triangular inverses + two GEMMs).
HD void luPhaseSolveDirect(Par P, int nc, const double* LU, const int* rowOf, const double* CD, const double* ev, double* Xt, double* Kt) {
  const int nut = NU - nc;
  for (int it = P.tid; it < NX + 1 + nut; it += P.nt) {
    double z[NC_MAX];
    if (it <= NX) {
#pragma unroll
      for (int i = 0; i < NC_MAX; ++i) z[i] = (i < nc) ? (it < NX ? CD[rowOf[i] + NC_MAX * it] : ev[rowOf[i]]) : 0.0;
#pragma unroll
      for (int i = 1; i < NC_MAX; ++i) {
        if (i < nc) {
          double s = z[i];
#pragma unroll
          for (int j = 0; j < i; ++j) s = fma(-LU[i + LU_LD * j], z[j], s);
          z[i] = s;
        }
      }
    } else {
      const int kk = it - NX - 1;
#pragma unroll
      for (int i = 0; i < NC_MAX; ++i) z[i] = (i < nc) ? LU[i + LU_LD * (nc + kk)] : 0.0;
    }
#pragma unroll
    for (int i = NC_MAX - 1; i >= 0; --i) {
      if (i < nc) {
        double s = z[i];
#pragma unroll
        for (int j = i + 1; j < NC_MAX; ++j)
          if (j < nc) s = fma(-LU[i + LU_LD * j], z[j], s);
        z[i] = s / LU[i + LU_LD * i];
      }
    }
    double* dst = (it <= NX) ? Xt + NC_MAX * it : Kt + NC_MAX * (it - NX - 1);
#pragma unroll
    for (int i = 0; i < NC_MAX; ++i)
      if (i < nc) dst[i] = -z[i];
  }
}

// scatter the solution through the column permutation to the layout the remap kernel reads (global memory):
//   Px (35 x 58), u0 (35), Pu (35 x nut)
HD void luPhaseInversePerm(Par P, const int* colOf, int* posOf) {
  for (int t = P.tid; t < NU; t += P.nt) posOf[colOf[t]] = t;
}
HD void luPhaseScatter(Par P, int nc, const int* posOf, const double* Xt, const double* Kt, double* Pu, double* Px, double* u0) {
  const int nut = NU - nc;
  for (int it = P.tid; it < NU * (NX + 1 + NUT_MAX); it += P.nt) {
    const int i = it % NU, j = it / NU;
    const int pos = posOf[i];
    if (j < NX) {
      Px[i + NU * j] = (pos < nc) ? Xt[pos + NC_MAX * j] : 0.0;
    } else if (j == NX) {
      u0[i] = (pos < nc) ? Xt[pos + NC_MAX * NX] : 0.0;
    } else {
      const int kk = j - NX - 1;
      Pu[i + NU * kk] = (kk < nut) ? ((pos < nc) ? Kt[pos + NC_MAX * kk] : (pos == nc + kk ? 1.0 : 0.0)) : 0.0;
    }
  }
}


// ---- K1b phase bodies shared by the host schedule (wb_node_b.inc) and the two CUDA schedules (wb_node_b1.inc / wb_node_b2.inc) -----------
// dense swing-foot rows: [Q S'; S R] = JS' JS (JS: nsw x 93, ld nsw, in shared memory or straight from the record in HBM/L2)
// qScale: factor on the contributions to Q alone (1 when Q is scaled with the other blocks on the way out; dt when Q accumulates in its
// final place in HBM/L2, wb_node_b2.inc)
template <class PAR>
HD void pjSwingRows(PAR P, int nsw, const double* __restrict__ JS, const PjWs& s, double qScale = 1.0) {
  par_mma_gemm<true, false, 4>(P, NX, NX, nsw, qScale, JS, nsw, JS, nsw, s.Q, NX);
  par_mma_gemm<true, false, 4>(rot(P, 128), NU, NX, nsw, 1.0, JS + nsw * NX, nsw, JS, nsw, s.S, NU);
  par_mma_gemm<true, false, 3>(rot(P, 64), NU, NU, nsw, 1.0, JS + nsw * NX, nsw, JS + nsw * NX, nsw, s.R, NU);
}
// structured rows: JU' JU on their 27-column support, added to the Hessian
HD void pjPhaseStructRows(Par P, int nru, const PjWs& s, double qScale = 1.0) {
  if (nru == 0) return;
  for (int it = P.tid; it < NUC * NUC; it += P.nt) {
    const int a = it % NUC, c = it / NUC;
    if (a < 15 && c >= 15) continue;   // that block is S' : only S is stored
    double v = 0.0;
    for (int k = 0; k < nru; ++k) v = fma(s.JU[k + JU_MAX * a], s.JU[k + JU_MAX * c], v);
    if (a < 15 && c < 15) s.Q[(3 + a) + NX * (3 + c)] += qScale * v;
    else if (a >= 15 && c < 15) s.S[(a - 15) + NU * (3 + c)] += v;
    else if (a >= 15 && c >= 15) s.R[(a - 15) + NU * (c - 15)] += v;
  }
}
// Hessian: diagonal (tracking cost, joint limits, curvature shift) and the friction-cone blocks of the stance feet.  Touches entries the
// structured rows also touch: never in the same phase as pjPhaseStructRows or a GEMM writing Q / R.
HD void pjPhaseHessianDiag(Par P, const double* __restrict__ mid, const PjWs& s, double qScale = 1.0) {
  for (int i = rot(P, 160).tid; i < NX; i += P.nt) s.Q[i + NX * i] += qScale * mid[Mid::HDG + i];
  for (int i = rot(P, 64).tid; i < NU + 18; i += P.nt) {   // one item per touched entry of R: 35 diagonal, 12 off-diagonal friction entries
    if (i < NU) {
      double v = mid[Mid::HDG + NX + i];
      if (i < 12 && i % 6 < 3) v += mid[Mid::FRIC + 9 * (i / 6) + 4 * (i % 6)];
      s.R[i + NU * i] += v;
    } else {
      const int e = i - NU, c = e / 9, a = (e % 9) / 3, b = e % 3;
      if (a != b) s.R[(6 * c + a) + NU * (6 * c + b)] += mid[Mid::FRIC + 9 * c + 3 * a + b];
    }
  }
}
// dynamics in the projected inputs: A~ = A + B Px, B~ = B Pu, b~ = b + B u0 -- dense rows from D12 = B1 [X | x0 | K], the others from the
// fixed pattern of the integrator rows
HD void pjPhaseDynamicsOut(Par P, int nc, int nut, double dt, const PjWs& s, const int* colOf, const int* posOf, const NodeOut& out) {
  for (int it = P.tid; it < NX * (NX + 1 + nut); it += P.nt) {
    const int r = it % NX, c = it / NX;
    const int dense = r < 6 ? r : ((r >= NV && r < NV + 6) ? 6 + r - NV : -1);
    double acc;
    if (dense >= 0) {
      acc = s.D12[dense + 12 * c];
      if (c < NX) acc += s.AB12[dense + 12 * c];
      else if (c == NX) acc += s.bvec[r];
      else acc += s.AB12[dense + 12 * (NX + colOf[nc + c - NX - 1])];
    } else {
      const int j = (r < NV) ? r - 6 : r - NV - 6;
      const double coef = (r < NV) ? 0.5 * dt * dt : dt;
      const int pos = posOf[12 + j];
      if (c <= NX) {
        acc = (pos < nc) ? coef * s.Xt[pos + NC_MAX * c] : 0.0;
        if (c < NX) acc += (c == r ? 1.0 : 0.0) + ((r < NV && c == NV + r) ? dt : 0.0);
        else acc += s.bvec[r];
      } else {
        const int kk = c - NX - 1;
        acc = (pos < nc) ? coef * s.Kt[pos + NC_MAX * kk] : (pos == nc + kk ? coef : 0.0);
      }
    }
    if (c < NX) out.A[r + NX * c] = acc;
    else if (c == NX) out.b[r] = acc;
    else out.Bt[r + NX * (c - NX - 1)] = acc;
  }
}
// gathers in pivot order: T11 = S1, T12 = S2, R11, R21, W = R12, V = R22, rr = r
HD void pjPhaseGather(Par P, int nc, const int* colOf, const PjWs& s) {
  for (int it = P.tid; it < NU * NX; it += P.nt) {
    const int t = it % NU, c = it / NU;
    const double v = s.S[colOf[t] + NU * c];
    if (t < nc) s.T11[t + NC_MAX * c] = v;
    else s.T12[(t - nc) + NUT_MAX * c] = v;
  }
  for (int it = P.tid; it < NU * NU; it += P.nt) {
    const int t = it % NU, u = it / NU;
    const double v = s.R[colOf[t] + NU * colOf[u]];
    if (t < nc) {
      if (u < nc) s.R11[t + NC_MAX * u] = v;
      else s.W[t + NC_MAX * (u - nc)] = v;
    } else {
      if (u < nc) s.R21[(t - nc) + NUT_MAX * u] = v;
      else s.V[(t - nc) + NUT_MAX * (u - nc)] = v;
    }
  }
  for (int t = P.tid; t < NU; t += P.nt) s.rr[t] = s.gq[NX + colOf[t]];
}
// projected cost blocks, scaled by dt
HD void pjPhaseOutputs(Par P, int nc, int nut, double dt, const PjWs& s, const NodeOut& out) {
  if (s.Q != out.Q)   // (Q accumulated in place, already scaled: wb_node_b2.inc)
    for (int i = P.tid; i < NX * NX; i += P.nt) out.Q[i] = dt * s.Q[i];
  for (int i = P.tid; i < NX; i += P.nt) out.q[i] = dt * s.gq[i];
  for (int i = P.tid; i < nut * NX; i += P.nt) out.St[(i % nut) + NUT_MAX * (i / nut)] = dt * s.T12[(i % nut) + NUT_MAX * (i / nut)];
  for (int i = P.tid; i < nut * nut; i += P.nt) out.Rt[(i % nut) + NUT_MAX * (i / nut)] = dt * s.V[(i % nut) + NUT_MAX * (i / nut)];
  for (int i = P.tid; i < nut; i += P.nt) out.rt[i] = dt * s.rr[nc + i];
  if (P.tid == 0) *out.nut = nut;
}
// optional raw (pre-projection) block dump, oracle layout; cost already multiplied by dt.  Part A: everything but the Hessian blocks.
HD void pjPhaseRawA(Par P, int nc, double dt, const double* __restrict__ mid, double* o) {
  for (int i = P.tid; i < NX * NZ; i += P.nt) o[i] = abEntry(mid + Mid::AB12, dt, i % NX, i / NX);
  o += NX * NZ;
  for (int i = P.tid; i < NX; i += P.nt) o[i] = mid[Mid::B_ + i];
  o += NX;
  o += NX * NX + NU * NX + NU * NU;   // Hessian blocks: pjPhaseRawB, once Q, S, R exist
  for (int i = P.tid; i < NX; i += P.nt) o[i] = dt * mid[Mid::GQ + i];
  o += NX;
  for (int i = P.tid; i < NU; i += P.nt) o[i] = dt * mid[Mid::GQ + NX + i];
  o += NU;
  if (P.tid == 0) o[0] = mid[Mid::META + 4];
  o += 1;
  for (int i = P.tid; i < NC_MAX * NZ; i += P.nt) o[i] = (i % NC_MAX < nc) ? mid[Mid::CD + i] : 0.0;
  o += NC_MAX * NZ;
  for (int i = P.tid; i < NC_MAX; i += P.nt) o[i] = (i < nc) ? mid[Mid::E + i] : 0.0;
  o += NC_MAX;
  if (P.tid == 0) o[0] = nc;
}
HD void pjPhaseRawB(Par P, double dt, const PjWs& s, double* raw, double qScale = 1.0) {
  double* o = raw + NX * NZ + NX;
  for (int i = P.tid; i < NX * NX; i += P.nt) o[i] = (dt / qScale) * s.Q[i];
  o += NX * NX;
  for (int i = P.tid; i < NU * NX; i += P.nt) o[i] = dt * s.S[i];
  o += NU * NX;
  for (int i = P.tid; i < NU * NU; i += P.nt) o[i] = dt * s.R[i];
}

}  // namespace b200sqp

namespace b200sqp {
// workspace of the value-only rollout kernel (K3)
struct RoWs {
  DynWs* dyn;
  double *fs, *xs, *FV, *FP, *pv, *valS, *ev, *sc, *xa, *xna, *ua;
  RowWs rw;
};
HD size_t roWsDoubles() {
  return DYN_VALUE_DOUBLES + 4 * NX + NX + 2 * FQ + 3 * NFRAMES + 96 + 2 * FQ + rowWsDoubles() + NC_MAX + 8 + 2 * NX + NU + 5;
}
HD void roWsMap(double* base, RoWs& r) {
  r.dyn = reinterpret_cast<DynWs*>(base);   // the value-only prefix: nothing behind DYN_VALUE_DOUBLES is touched by K3's phases
  r.fs = base + DYN_VALUE_DOUBLES;
  r.xs = r.fs + 4 * NX;
  r.FV = r.xs + NX;
  r.FP = r.FV + 2 * FQ;
  r.pv = r.FP + 3 * NFRAMES;
  r.valS = r.pv + 96;
  rowWsMap(r.valS + 2 * FQ, r.rw);
  r.ev = r.valS + 2 * FQ + rowWsDoubles();
  r.sc = r.ev + NC_MAX;
  r.xa = r.sc + 8;
  r.xna = r.xa + NX;
  r.ua = r.xna + NX;
}
}  // namespace b200sqp
