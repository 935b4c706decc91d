// Kernels of the centroidal SQP iteration.  The centroidal OCP (nx = nu = 35) runs inside the whole-body solver's device state: every
// per-node array keeps the whole-body strides (state 58, projected input 23) with the 23 surplus state slots as inert dummy states
// (x = 0, A = 1 on their diagonal, zero cost, zero rows of B and columns of S / Px), so K2, the remap / line-search / acceptance kernels and
// the host loop are shared unchanged.  The padding never changes a result on the live block (only exact zeros are added); it costs
// (58/35)^3 in K2 flops, irrelevant for the single-instance correctness configurations this path serves (BASELINE configs 0-1).
//   cen_lq_kernel      one CTA per (instance, node): LQ approximation (cen_ocp.cuh) -> raw stage block (oracle layout, nx = nu = 35)
//   cen_proj_kernel    one CTA per (instance, stage): luConstraintProjection + changeOfInputVariables -> padded QP stage record
//   cen_rollout_kernel one CTA per (instance, node): value-only defect / cost / constraint norms at x + alpha dx
#pragma once
#include "cen_ocp.cuh"

namespace b200sqp {

constexpr long long kCenRawPer = 2LL * CNX * CNX + 2LL * CNX * CNU + CNU * CNU + 2 * CNX + CNU + 1 + 14LL * (CNX + CNU + 1) + 1;
constexpr int CEN_PROJ_THREADS = 128;

struct CenDevModel {
  CenOcpModel ocp;
  double QfdPad[NX];  // final-cost weights on the padded state
};

struct CenLqWs {
  double *J;      // [CEN_ROWS][CNZ] row-major: Jacobian rows of the Gauss-Newton residuals (0..49) and the penalty constraints (50..73)
  double *val, *w2, *gc;  // per row: value, curvature weight of J'J, gradient coefficient
  double *Jg, *gval;      // equality constraints: [NC_MAX x CNZ] column-major (ld NC_MAX), values
  double *AB, *xplus;     // [CNX x CNZ] column-major, RK4 image
  double *H, *grad;       // [CNZ x CNZ] column-major, gradient
  double *tref, *sc;      // task-space reference (13), scalars {cost, fricD1[2]}
};
HD size_t cenLqWsDoubles() { return CEN_ROWS * CNZ + 3 * CEN_ROWS + NC_MAX * CNZ + NC_MAX + CNX * CNZ + CNX + CNZ * CNZ + CNZ + 13 + 8; }
HD void cenLqWsMap(double* b, CenLqWs& s) {
  s.J = b;
  s.val = s.J + CEN_ROWS * CNZ;
  s.w2 = s.val + CEN_ROWS;
  s.gc = s.w2 + CEN_ROWS;
  s.Jg = s.gc + CEN_ROWS;
  s.gval = s.Jg + NC_MAX * CNZ;
  s.AB = s.gval + NC_MAX;
  s.xplus = s.AB + CNX * CNZ;
  s.H = s.xplus + CNX;
  s.grad = s.H + CNZ * CNZ;
  s.tref = s.grad + CNZ;
  s.sc = s.tref + 13;
}

// thread `dir` < 70 deposits its tangents, thread 0 also the values
HD void cenPhaseDeposit(int tid, const NodeIn& n, const CenDirOut& o, CenLqWs s) {
  const int nres = cenResidualRows(n), nc = cenConstraintRows(n);
  if (tid < CNZ) {
    for (int r = 0; r < CEN_MAX_RES; ++r) s.J[r * CNZ + tid] = (r < nres) ? o.res[r].d : 0.0;
    for (int r = 0; r < CEN_PEN_ROWS; ++r) s.J[(CEN_MAX_RES + r) * CNZ + tid] = o.pen[r].d;
    for (int r = 0; r < NC_MAX; ++r) s.Jg[r + NC_MAX * tid] = (r < nc) ? o.g[r].d : 0.0;
    for (int i = 0; i < CNX; ++i) s.AB[i + CNX * tid] = o.xplus[i].d;
  }
  if (tid == 0) {
    for (int r = 0; r < CEN_MAX_RES; ++r) s.val[r] = (r < nres) ? o.res[r].v : 0.0;
    for (int r = 0; r < CEN_PEN_ROWS; ++r) s.val[CEN_MAX_RES + r] = o.pen[r].v;
    for (int r = 0; r < NC_MAX; ++r) s.gval[r] = (r < nc) ? o.g[r].v : 0.0;
    for (int i = 0; i < CNX; ++i) s.xplus[i] = o.xplus[i].v;
  }
}

// activity of penalty row r (0..23): contact-moment rows need their foot in stance, collision rows are off in double stance
HD bool cenPenaltyActive(const NodeIn& n, int r) { return r < 8 ? n.contact[r / 4] != 0 : !(n.contact[0] && n.contact[1]); }

// row scalars + the node's cost value (one thread; 74 rows).  Returns the cost (without dt).
HD double cenRowScalars(const CenOcpModel& m, const NodeIn& n, const double* val, double* w2, double* gc, double* fricD1) {
  double cost = 0.0;
  const int nres = cenResidualRows(n);
  for (int r = 0; r < CEN_MAX_RES; ++r) {
    const bool on = r < nres;
    if (w2) {
      w2[r] = on ? 1.0 : 0.0;
      gc[r] = on ? val[r] : 0.0;
    }
    if (on) cost += 0.5 * val[r] * val[r];
  }
  for (int r = 0; r < CEN_PEN_ROWS; ++r) {
    double v = 0.0, d1 = 0.0, d2 = 0.0;
    if (cenPenaltyActive(n, r)) {
      if (r < 8) penRelaxed(m.momMu, m.momDelta, val[CEN_MAX_RES + r], v, d1, d2);
      else penPwPoly(m.collMu, m.collDelta, val[CEN_MAX_RES + r], v, d1, d2);
    }
    if (w2) {
      w2[CEN_MAX_RES + r] = d2;
      gc[CEN_MAX_RES + r] = d1;
    }
    cost += v;
  }
  // analytic terms: quadratic tracking, joint limits, friction cone
  const double yaw = n.x[9];
  const double gcyc = n.armPhase * (cos(yaw) * n.xref[0] + sin(yaw) * n.xref[1]);
  for (int i = 0; i < CNX; ++i) {
    double xn = n.xref[i];
    if (i >= 12) {
      const int j = i - 12;
      if (j == m.armJoint[0] || j == m.armJoint[2]) xn += -0.15 * gcyc;
      if (j == m.armJoint[1] || j == m.armJoint[3]) xn += 0.15 * gcyc;
      double v, d1, d2;
      penPwPoly(m.jlMu, m.jlDelta, m.qhi[j] - n.x[i], v, d1, d2);
      cost += v;
      penPwPoly(m.jlMu, m.jlDelta, n.x[i] - m.qlo[j], v, d1, d2);
      cost += v;
    }
    const double dx = n.x[i] - xn;
    cost += 0.5 * m.Qd[i] * dx * dx;
  }
  const int ns = n.contact[0] + n.contact[1];
  const double fz = ns > 0 ? m.kin.mtot * 9.81 / ns : 0.0;
  for (int i = 0; i < CNU; ++i) {
    const double un = (i < 12 && i % 6 == 2 && n.contact[i / 6]) ? fz : 0.0;
    const double du = n.u[i] - un;
    cost += 0.5 * m.Rd[i] * du * du;
  }
  for (int c = 0; c < 2; ++c) {
    double d1 = 0.0;
    if (n.contact[c]) {
      const double* F = n.u + 6 * c;
      const double h = m.fricCoeff * F[2] - sqrt(F[0] * F[0] + F[1] * F[1] + m.fricReg);
      double v, d2;
      penRelaxed(m.fricMu, m.fricDelta, h, v, d1, d2);
      cost += v;
    }
    if (fricD1) fricD1[c] = d1;
  }
  return cost;
}

// H = sum_r w2_r J_r' J_r + analytic blocks; grad = sum_r gc_r J_r + analytic gradient   (all threads)
HD void cenPhaseHessian(Par P, const CenOcpModel& m, const NodeIn& n, CenLqWs s) {
  const double shift = -(s.sc[1] + s.sc[2]) * m.fricShift;
  for (int it = P.tid; it < CNZ * CNZ; it += P.nt) {
    const int i = it % CNZ, j = it / CNZ;
    double a = 0.0;
    for (int r = 0; r < CEN_ROWS; ++r) a = fma(s.w2[r] * s.J[r * CNZ + i], s.J[r * CNZ + j], a);
    if (i == j) {
      a += (i < CNX) ? m.Qd[i] : m.Rd[i - CNX];
      a += shift;
      if (i >= 12 && i < CNX) {
        const int jj = i - 12;
        double v, d1, d2;
        penPwPoly(m.jlMu, m.jlDelta, m.qhi[jj] - n.x[i], v, d1, d2);
        a += d2;
        penPwPoly(m.jlMu, m.jlDelta, n.x[i] - m.qlo[jj], v, d1, d2);
        a += d2;
      }
    }
    // friction cone (quadratic-order constraint): p'' dh dh' + p' ddh on the force block of a stance foot
    const int ui = i - CNX, uj = j - CNX;
    if (ui >= 0 && uj >= 0 && ui < 12 && uj < 12 && ui / 6 == uj / 6 && ui % 6 < 3 && uj % 6 < 3 && n.contact[ui / 6]) {
      const int c = ui / 6, a3 = ui % 6, b3 = uj % 6;
      const double* F = n.u + 6 * c;
      const double Ft2 = F[0] * F[0] + F[1] * F[1] + m.fricReg, Ft = sqrt(Ft2), Ft32 = Ft * Ft2;
      const double h = m.fricCoeff * F[2] - Ft;
      double v, d1, d2;
      penRelaxed(m.fricMu, m.fricDelta, h, v, d1, d2);
      const double dh[3] = {-F[0] / Ft, -F[1] / Ft, m.fricCoeff};
      double ddh = 0.0;
      if (a3 < 2 && b3 < 2) ddh = (a3 == b3) ? -(F[1 - a3] * F[1 - a3] + m.fricReg) / Ft32 : F[0] * F[1] / Ft32;
      a += d2 * dh[a3] * dh[b3] + d1 * ddh;
    }
    s.H[it] = a;
  }
  const double yaw = n.x[9];
  const double gcyc = n.armPhase * (cos(yaw) * n.xref[0] + sin(yaw) * n.xref[1]);
  const int ns = n.contact[0] + n.contact[1];
  const double fz = ns > 0 ? m.kin.mtot * 9.81 / ns : 0.0;
  for (int i = P.tid; i < CNZ; i += P.nt) {
    double g = 0.0;
    for (int r = 0; r < CEN_ROWS; ++r) g = fma(s.gc[r], s.J[r * CNZ + i], g);
    if (i < CNX) {
      double xn = n.xref[i];
      if (i >= 12) {
        const int j = i - 12;
        if (j == m.armJoint[0] || j == m.armJoint[2]) xn += -0.15 * gcyc;
        if (j == m.armJoint[1] || j == m.armJoint[3]) xn += 0.15 * gcyc;
        double v, d1u, d1l, d2;
        penPwPoly(m.jlMu, m.jlDelta, m.qhi[j] - n.x[i], v, d1u, d2);
        penPwPoly(m.jlMu, m.jlDelta, n.x[i] - m.qlo[j], v, d1l, d2);
        g += d1l - d1u;
      }
      g += m.Qd[i] * (n.x[i] - xn);
    } else {
      const int ui = i - CNX;
      const double un = (ui < 12 && ui % 6 == 2 && n.contact[ui / 6]) ? fz : 0.0;
      g += m.Rd[ui] * (n.u[ui] - un);
      if (ui < 12 && ui % 6 < 3 && n.contact[ui / 6]) {
        const double* F = n.u + 6 * (ui / 6);
        const double Ft = sqrt(F[0] * F[0] + F[1] * F[1] + m.fricReg);
        const double dh = (ui % 6 < 2) ? -F[ui % 6] / Ft : m.fricCoeff;
        g += s.sc[1 + ui / 6] * dh;
      }
    }
    s.grad[i] = g;
  }
}

// raw stage block in the oracle layout (cost scaled by dt as in setupIntermediateNode, Transcription.cpp:40-94)
HD void cenPhaseRaw(Par P, const NodeIn& n, CenLqWs s, double* raw, double* perf) {
  const double dt = n.dt;
  const int nc = cenConstraintRows(n);
  double* o = raw;
  for (int i = P.tid; i < CNX * CNZ; i += P.nt) o[i] = s.AB[i];   // A then B, both column-major with ld 35
  o += CNX * CNZ;
  for (int i = P.tid; i < CNX; i += P.nt) o[i] = s.xplus[i] - n.xnext[i];
  o += CNX;
  for (int it = P.tid; it < CNX * CNX; it += P.nt) o[it] = dt * s.H[(it % CNX) + CNZ * (it / CNX)];
  o += CNX * CNX;
  for (int it = P.tid; it < CNU * CNX; it += P.nt) o[it] = dt * s.H[(CNX + it % CNU) + CNZ * (it / CNU)];   // S = dfdux (nu x nx)
  o += CNU * CNX;
  for (int it = P.tid; it < CNU * CNU; it += P.nt) o[it] = dt * s.H[(CNX + it % CNU) + CNZ * (CNX + it / CNU)];
  o += CNU * CNU;
  for (int i = P.tid; i < CNZ; i += P.nt) o[i] = dt * s.grad[i];   // q then r
  o += CNZ;
  if (P.tid == 0) o[0] = dt * s.sc[0];
  o += 1;
  for (int it = P.tid; it < NC_MAX * CNZ; it += P.nt) o[it] = s.Jg[it];   // C (14 x 35) then D (14 x 35), ld 14
  o += NC_MAX * CNZ;
  for (int i = P.tid; i < NC_MAX; i += P.nt) o[i] = s.gval[i];
  o += NC_MAX;
  if (P.tid == 0) {
    o[0] = nc;
    double dsse = 0.0, esse = 0.0;
    for (int i = 0; i < CNX; ++i) {
      const double df = s.xplus[i] - n.xnext[i];
      dsse = fma(df, df, dsse);
    }
    for (int i = 0; i < nc; ++i) esse = fma(s.gval[i], s.gval[i], esse);
    perf[0] = dt * s.sc[0];
    perf[1] = dt * dsse;
    perf[2] = dt * esse;
  }
}

// ---- projection + change of input variables from the raw block (cen_proj_kernel) -------------------------------------------------------------
struct CenPjWs {
  double *A, *B, *b, *Q, *S, *R, *q, *r;   // compact copies of the raw block (ld 35)
  double *CD, *ev, *LU, *Xt, *Kt;          // LU workspaces in the whole-body shapes (C padded to 58 columns)
  double *Px, *Pu, *u0;                    // compact projection (35 x 35, 35 x 23, 35)
  double *T, *RP, *rr;                     // S + R Px ; R Px / R Pu ; r + R u0
  int* iw;
};
HD size_t cenPjWsDoubles() {
  return 2 * CNX * CNX + CNX + CNX * CNX + 2 * CNU * CNX + CNX + CNU + NC_MAX * NX + NC_MAX + LU_LD * NU + NC_MAX * (NX + 1) + NC_MAX * NUT_MAX +
         CNU * CNX + CNU * NUT_MAX + CNU + 2 * CNU * CNX + CNU + 64;
}
HD void cenPjWsMap(double* b, CenPjWs& s) {
  s.A = b;
  s.B = s.A + CNX * CNX;
  s.b = s.B + CNX * CNU;
  s.Q = s.b + CNX;
  s.S = s.Q + CNX * CNX;
  s.R = s.S + CNU * CNX;
  s.q = s.R + CNU * CNU;
  s.r = s.q + CNX;
  s.CD = s.r + CNU;
  s.ev = s.CD + NC_MAX * NX;
  s.LU = s.ev + NC_MAX;
  s.Xt = s.LU + LU_LD * NU;
  s.Kt = s.Xt + NC_MAX * (NX + 1);
  s.Px = s.Kt + NC_MAX * NUT_MAX;
  s.Pu = s.Px + CNU * CNX;
  s.u0 = s.Pu + CNU * NUT_MAX;
  s.T = s.u0 + CNU;
  s.RP = s.T + CNU * CNX;
  s.rr = s.RP + CNU * CNX;
  s.iw = reinterpret_cast<int*>(s.rr + CNU);
}

}  // namespace b200sqp
