// ORACLE (test infrastructure only) -- C entry points for the Python tests (ctypes) and bench.py's CPU-baseline leg.
// Product code never links this library.
#include <atomic>
#include <cstring>
#include <memory>

#include "sqp.hpp"
#include "test_problems.hpp"
#include "trajectory_spreading.hpp"
#ifdef ORC_WITH_WB
#include "wb_problem.hpp"
#include "cen_dynamics.hpp"
#include "cen_problem.hpp"
#endif

using namespace orc;

namespace {
Mat loadMat(const double* p, int r, int c, int ld = -1) {
  if (ld < 0) ld = r;
  Mat m(r, c);
  for (int j = 0; j < c; ++j)
    for (int i = 0; i < r; ++i) m(i, j) = p[static_cast<size_t>(j) * ld + i];
  return m;
}
void storeMat(const Mat& m, double* p, int ld = -1) {
  if (ld < 0) ld = m.r;
  for (int j = 0; j < m.c; ++j)
    for (int i = 0; i < m.r; ++i) p[static_cast<size_t>(j) * ld + i] = m(i, j);
}
Vec loadVec(const double* p, int n) { return Vec(p, p + n); }
void storeLog(const std::vector<IterationLog>& log, double* out, int cap, int* nIter) {
  // per iteration 16 doubles: baseline{merit,cost,dynSSE,eqSSE}, after{merit,cost,dynSSE,eqSSE}, stepSize, stepType, dx_norm,
  // du_norm, armijo, convergence, 0, 0
  *nIter = static_cast<int>(log.size());
  for (int i = 0; i < *nIter && i < cap; ++i) {
    double* o = out + 16 * i;
    const auto& e = log[i];
    o[0] = e.baseline.merit;
    o[1] = e.baseline.cost;
    o[2] = e.baseline.dynamicsViolationSSE;
    o[3] = e.baseline.equalityConstraintsSSE;
    o[4] = e.step.performanceAfterStep.merit;
    o[5] = e.step.performanceAfterStep.cost;
    o[6] = e.step.performanceAfterStep.dynamicsViolationSSE;
    o[7] = e.step.performanceAfterStep.equalityConstraintsSSE;
    o[8] = e.step.stepSize;
    o[9] = static_cast<double>(static_cast<int>(e.step.stepType));
    o[10] = e.step.dx_norm;
    o[11] = e.step.du_norm;
    o[12] = e.armijoDescentMetric;
    o[13] = static_cast<double>(static_cast<int>(e.convergence));
    o[14] = o[15] = 0.0;
  }
}
}  // namespace

extern "C" {

// trajectorySpread on flat arrays, in place: t [n], x [n][nx], tags [n] (an integer tag per sample, e.g. its mode), events [n_events_in] (one value per
// event of the old trajectory, filtered with extractEventsArray).  Returns the new length; flags = {willTruncate, willSpread};
// post [<= n] receives the updated post-event indices (n_post), events_out the kept event data (n_events_out).
int orc_trajectory_spread(int n_old_ev, const double* old_ev, const int* old_modes, int n_new_ev, const double* new_ev, const int* new_modes, int n, int nx,
                          double* t, double* x, int* tags, int* flags, int* post, int* n_post, int n_events_in, const double* events_in,
                          double* events_out, int* n_events_out) {
  oracle::SpreadSchedule o{std::vector<double>(old_ev, old_ev + n_old_ev), std::vector<int>(old_modes, old_modes + n_old_ev + 1)};
  oracle::SpreadSchedule w{std::vector<double>(new_ev, new_ev + n_new_ev), std::vector<int>(new_modes, new_modes + n_new_ev + 1)};
  std::vector<double> tv(t, t + n), xv(x, x + static_cast<size_t>(n) * nx), gv(tags, tags + n);
  const oracle::SpreadPlan plan = oracle::spreadPlan(o, w, tv);
  const size_t m = oracle::spreadApply(plan, xv, nx);
  oracle::spreadApply(plan, gv, 1);
  oracle::spreadApplyTime(plan, tv);
  std::copy(tv.begin(), tv.end(), t);
  std::copy(xv.begin(), xv.end(), x);
  for (size_t i = 0; i < m; ++i) tags[i] = static_cast<int>(gv[i]);
  flags[0] = plan.willTruncate;
  flags[1] = plan.willSpread;
  *n_post = static_cast<int>(plan.postEventIndices.size());
  for (size_t i = 0; i < plan.postEventIndices.size(); ++i) post[i] = static_cast<int>(plan.postEventIndices[i]);
  int ne = 0;
  for (size_t i = plan.keepEventsFirst; i < plan.keepEventsLast && static_cast<int>(i) < n_events_in; ++i) events_out[ne++] = events_in[i];
  *n_events_out = ne;
  return static_cast<int>(m);
}

int orc_time_discretization(double t0, double tf, double dt, const double* ev, int nev, double* t_out, int* ev_out, int cap) {
  auto td = timeDiscretizationWithEvents(t0, tf, dt, std::vector<double>(ev, ev + nev));
  const int n = static_cast<int>(td.size());
  for (int i = 0; i < n && i < cap; ++i) {
    t_out[i] = td[i].time;
    ev_out[i] = static_cast<int>(td[i].event);
  }
  return n;
}

// Stage arrays are padded to numax inputs; stage k uses the leading nu[k] columns/rows.
// A[k]: nx x nx, B[k]: nx x numax, S[k]: numax x nx (ld numax), R[k]: numax x numax, K[k]: numax x nx (ld numax)
int orc_riccati(int N, int nx, int numax, const int* nu, const double* A, const double* B, const double* b, const double* Q,
                const double* S, const double* R, const double* q, const double* r, const double* dx0, double reg, double* dx,
                double* du, double* P, double* p, double* K, double* kff) {
  std::vector<LinApprox> dyn(N);
  std::vector<QuadApprox> cost(N + 1);
  for (int k = 0; k < N; ++k) {
    dyn[k].dfdx = loadMat(A + static_cast<size_t>(k) * nx * nx, nx, nx);
    dyn[k].dfdu = loadMat(B + static_cast<size_t>(k) * nx * numax, nx, nu[k]);
    dyn[k].f = loadVec(b + static_cast<size_t>(k) * nx, nx);
    cost[k].dfdxx = loadMat(Q + static_cast<size_t>(k) * nx * nx, nx, nx);
    cost[k].dfdux = loadMat(S + static_cast<size_t>(k) * numax * nx, nu[k], nx, numax);
    cost[k].dfduu = loadMat(R + static_cast<size_t>(k) * numax * numax, nu[k], nu[k], numax);
    cost[k].dfdx = loadVec(q + static_cast<size_t>(k) * nx, nx);
    cost[k].dfdu = loadVec(r + static_cast<size_t>(k) * numax, nu[k]);
  }
  cost[N].dfdxx = loadMat(Q + static_cast<size_t>(N) * nx * nx, nx, nx);
  cost[N].dfdx = loadVec(q + static_cast<size_t>(N) * nx, nx);
  RiccatiSolution s = solveRiccati(loadVec(dx0, nx), dyn, cost, reg);
  if (!s.ok) return -1;
  for (int k = 0; k <= N; ++k) {
    std::memcpy(dx + static_cast<size_t>(k) * nx, s.dx[k].data(), sizeof(double) * nx);
    if (P) storeMat(s.P[k], P + static_cast<size_t>(k) * nx * nx);
    if (p) std::memcpy(p + static_cast<size_t>(k) * nx, s.p[k].data(), sizeof(double) * nx);
  }
  for (int k = 0; k < N; ++k) {
    std::memset(du + static_cast<size_t>(k) * numax, 0, sizeof(double) * numax);
    if (nu[k]) std::memcpy(du + static_cast<size_t>(k) * numax, s.du[k].data(), sizeof(double) * nu[k]);
    if (K) {
      std::memset(K + static_cast<size_t>(k) * numax * nx, 0, sizeof(double) * numax * nx);
      if (nu[k]) storeMat(s.K[k], K + static_cast<size_t>(k) * numax * nx, numax);
    }
    if (kff) {
      std::memset(kff + static_cast<size_t>(k) * numax, 0, sizeof(double) * numax);
      if (nu[k]) std::memcpy(kff + static_cast<size_t>(k) * numax, s.kff[k].data(), sizeof(double) * nu[k]);
    }
  }
  return 0;
}

// luConstraintProjection: Pu (nu x (nu-rank)), Px (nu x nx), u0 (nu); returns rank
int orc_lu_projection(int nc, int nx, int nu, const double* C, const double* D, const double* e, double* Pu, double* Px, double* u0) {
  LinApprox g;
  g.dfdx = loadMat(C, nc, nx);
  g.dfdu = loadMat(D, nc, nu);
  g.f = loadVec(e, nc);
  FullPivLU lu(g.dfdu);
  LinApprox pr = luConstraintProjection(g);
  storeMat(pr.dfdu, Pu);
  storeMat(pr.dfdx, Px);
  std::memcpy(u0, pr.f.data(), sizeof(double) * nu);
  return lu.rank;
}

// changeOfInputVariables on a cost and a dynamics block (in place). S is nu x nx.
void orc_change_of_input_variables(int nx, int nu, int nut, double* A, double* B, double* b, double* Q, double* S, double* R, double* q,
                                   double* r, double* c, const double* Pu, const double* Px, const double* u0, double* Bt, double* St,
                                   double* Rt, double* rt) {
  LinApprox d;
  d.dfdx = loadMat(A, nx, nx);
  d.dfdu = loadMat(B, nx, nu);
  d.f = loadVec(b, nx);
  QuadApprox cq;
  cq.dfdxx = loadMat(Q, nx, nx);
  cq.dfdux = loadMat(S, nu, nx);
  cq.dfduu = loadMat(R, nu, nu);
  cq.dfdx = loadVec(q, nx);
  cq.dfdu = loadVec(r, nu);
  cq.f = *c;
  const Mat mPu = loadMat(Pu, nu, nut), mPx = loadMat(Px, nu, nx);
  const Vec vu0 = loadVec(u0, nu);
  changeOfInputVariables(d, mPu, mPx, vu0);
  changeOfInputVariables(cq, mPu, mPx, vu0);
  storeMat(d.dfdx, A);
  storeMat(d.dfdu, Bt);
  std::memcpy(b, d.f.data(), sizeof(double) * nx);
  storeMat(cq.dfdxx, Q);
  storeMat(cq.dfdux, St);
  storeMat(cq.dfduu, Rt);
  std::memcpy(q, cq.dfdx.data(), sizeof(double) * nx);
  std::memcpy(rt, cq.dfdu.data(), sizeof(double) * nut);
  *c = cq.f;
}

// RK4 sensitivity of a linear system x' = A x + B u  (CORE/test/integration/testSensitivityIntegrator.cpp recipe)
void orc_rk4_sensitivity_linear(int nx, int nu, const double* A, const double* B, const double* x, const double* u, double dt, double* Ad,
                                double* Bd, double* xn, double* xn_value_only) {
  LinearQuadraticOcp ocp;
  ocp.nx = nx;
  ocp.nu = nu;
  ocp.A = loadMat(A, nx, nx);
  ocp.B = loadMat(B, nx, nu);
  LinApprox l = rk4SensitivityDiscretization(ocp, 0, 0.0, loadVec(x, nx), loadVec(u, nu), dt);
  storeMat(l.dfdx, Ad);
  storeMat(l.dfdu, Bd);
  std::memcpy(xn, l.f.data(), sizeof(double) * nx);
  Vec v = rk4Discretization(ocp, 0, 0.0, loadVec(x, nx), loadVec(u, nu), dt);
  std::memcpy(xn_value_only, v.data(), sizeof(double) * nx);
}

struct orc_lq_problem {
  int kind;  // 0 linear-quadratic, 1 circular kinematics
  int nx, nu;
  const double *A, *B, *G;  // G may be null
  const double *Q, *R, *P, *Qf, *Qe;  // Qe may be null
  const double *xRef, *uRef;
  int nModes;  // number of constraint rows sets (each 1 x ...): Cm (nModes x nx), Dm (nModes x nu), em (nModes)
  const double *Cm, *Dm, *em;
  int nEvents;
  const double* eventTimes;
  const int* modeSequence;  // nEvents+1 entries: constraint set index (or -1) active in each phase
  double t0, tf;
  const double* x0;
  double dt;
  int sqpIteration;
};

// Runs SqpOracle on a test problem from a DefaultInitializer guess (u = 0, x_k = x0).
// outputs: times/events (cap nodes), x (cap x nx), u (cap x nu), K (cap x nu x nx col-major per node, remapped), log (capIter x 16)
int orc_sqp_test_problem(const orc_lq_problem* pr, int cap, double* times, int* events, double* x, double* u, double* K, int capIter,
                         double* log, int* nIter) {
  std::unique_ptr<Ocp> ocp;
  LinearQuadraticOcp* lq = nullptr;
  if (pr->kind == 0) {
    auto o = std::make_unique<LinearQuadraticOcp>();
    o->nx = pr->nx;
    o->nu = pr->nu;
    o->A = loadMat(pr->A, pr->nx, pr->nx);
    o->B = loadMat(pr->B, pr->nx, pr->nu);
    if (pr->G) o->G = loadMat(pr->G, pr->nx, pr->nx);
    o->Q = loadMat(pr->Q, pr->nx, pr->nx);
    o->R = loadMat(pr->R, pr->nu, pr->nu);
    o->P = loadMat(pr->P, pr->nu, pr->nx);
    o->Qf = loadMat(pr->Qf, pr->nx, pr->nx);
    if (pr->Qe) {
      o->Qe = loadMat(pr->Qe, pr->nx, pr->nx);
      o->hasEventCost = true;
    }
    o->xRef = loadVec(pr->xRef, pr->nx);
    o->uRef = loadVec(pr->uRef, pr->nu);
    for (int m = 0; m < pr->nModes; ++m) {
      o->Cm.push_back(loadMat(pr->Cm + static_cast<size_t>(m) * pr->nx, 1, pr->nx));
      o->Dm.push_back(loadMat(pr->Dm + static_cast<size_t>(m) * pr->nu, 1, pr->nu));
      o->em.push_back({pr->em[m]});
    }
    lq = o.get();
    ocp = std::move(o);
  } else {
    ocp = std::make_unique<CircularKinematicsOcp>();
  }
  const std::vector<double> ev(pr->eventTimes, pr->eventTimes + pr->nEvents);
  auto td = timeDiscretizationWithEvents(pr->t0, pr->tf, pr->dt, ev);
  const int n = static_cast<int>(td.size());
  if (n > cap) return -n;
  if (lq && pr->nModes > 0) {
    lq->nodeMode.resize(n);
    for (int i = 0; i < n; ++i) {
      // ModeSchedule::modeAtTime(t): index = lower_bound(eventTimes, t)   (ocs2_core/src/reference/ModeSchedule.cpp)
      const double t = getIntervalStart(td[i]);
      lq->nodeMode[i] = pr->modeSequence[findIndexInTimeArray(ev, t)];
    }
  }
  SqpSettings s;
  s.dt = pr->dt;
  s.sqpIteration = pr->sqpIteration;
  SqpOracle solver(*ocp, s);
  const Vec x0 = loadVec(pr->x0, ocp->nx);
  std::vector<Vec> xs(n, x0), us(n - 1, vzero(ocp->nu));
  for (int i = 0; i + 1 < n; ++i)
    if (td[i].event == Event::PreEvent) us[i] = Vec();
  solver.run(td, x0, xs, us);
  for (int i = 0; i < n; ++i) {
    times[i] = td[i].time;
    events[i] = static_cast<int>(td[i].event);
    std::memcpy(x + static_cast<size_t>(i) * ocp->nx, xs[i].data(), sizeof(double) * ocp->nx);
    if (i < n - 1) {
      double* ui = u + static_cast<size_t>(i) * ocp->nu;
      std::memset(ui, 0, sizeof(double) * ocp->nu);
      if (!us[i].empty()) std::memcpy(ui, us[i].data(), sizeof(double) * ocp->nu);
      double* Ki = K + static_cast<size_t>(i) * ocp->nu * ocp->nx;
      std::memset(Ki, 0, sizeof(double) * ocp->nu * ocp->nx);
      if (solver.Kgain[i].r > 0) storeMat(solver.Kgain[i], Ki);
    }
  }
  storeLog(solver.log, log, capIter, nIter);
  return n;
}

}  // extern "C"

#ifdef ORC_WITH_WB
#include "wb_capi.inc"
#endif
