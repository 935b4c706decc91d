// ORACLE (test infrastructure only) -- CPU restatement of the multiple-shooting SQP iteration of ocs2::SqpSolver.
//
// Follows, function by function (paths relative to /root/reference/lib/ocs2_ros2/):
//   SqpSolver::runImpl / setupQuadraticSubproblem / getOCPSolution / takeStep / checkConvergence
//       ocs2_sqp/ocs2_sqp/src/SqpSolver.cpp:193-602
//   timeDiscretizationWithEvents                     ocs2_oc/src/oc_data/TimeDiscretization.cpp:40-114
//   rk4Discretization / rk4SensitivityDiscretization ocs2_core/src/integration/SensitivityIntegratorImpl.cpp:109-169
//   setupIntermediateNode / projectTranscription / setupEventNode / setupTerminalNode
//       ocs2_oc/src/multiple_shooting/Transcription.cpp:40-192
//   changeOfInputVariables                           ocs2_oc/src/approximate_model/ChangeOfInputVariables.cpp:34-108
//   luConstraintProjection                           ocs2_core/src/misc/LinearAlgebra.cpp:183-199
//   PerformanceIndex bookkeeping                     ocs2_oc/src/multiple_shooting/PerformanceIndexComputation.cpp:40-80,
//                                                    ocs2_oc/src/oc_data/PerformanceIndex.cpp:83-105
//   FilterLinesearch::acceptStep, armijoDescentMetric ocs2_oc/src/search_strategy/FilterLinesearch.cpp:34-89
//   remapProjectedInput / remapProjectedGain         ocs2_oc/src/multiple_shooting/Helpers.cpp:38-58
//   HpipmInterface::solve (x0 elimination) / getRiccatiFeedback / getRiccatiCostToGo
//       ocs2_sqp/hpipm_catkin/src/HpipmInterface.cpp:166-301,330-455
// HPIPM itself (giaf/hpipm@255ffdf, un-vendored) is restated as the classical Riccati recursion it runs for a
// QP without inequality constraints (d_ocp_qp_fact_solve_kkt_unconstr): per stage, Cholesky of
// [R+B'PB, .; S'+A'PB, Q+A'PA] + reg_prim*I, P = Q~ - Ls Ls', forward substitution for (dx,du).  The
// textbook recursion in hpipm_catkin/test/testHpipmInterface.cpp:281-304 pins P,p,K,k at 1e-9.
#pragma once
#include <array>
#include <cstdint>
#include <functional>
#include <limits>

#include "linalg.hpp"

namespace orc {

struct LinApprox {  // ocs2::VectorFunctionLinearApproximation (ocs2_core/include/ocs2_core/Types.h:145-157)
  Mat dfdx, dfdu;
  Vec f;
};
struct QuadApprox {  // ocs2::ScalarFunctionQuadraticApproximation (Types.h:234-240); dfdux is nu x nx
  Mat dfdxx, dfdux, dfduu;
  Vec dfdx, dfdu;
  double f = 0.0;
};

struct PerformanceIndex {  // ocs2_oc/include/ocs2_oc/oc_data/PerformanceIndex.h:42-98
  double merit = 0, cost = 0, dualFeasibilitiesSSE = 0, dynamicsViolationSSE = 0, equalityConstraintsSSE = 0,
         inequalityConstraintsSSE = 0, equalityLagrangian = 0, inequalityLagrangian = 0;
  PerformanceIndex& operator+=(const PerformanceIndex& o) {
    merit += o.merit;
    cost += o.cost;
    dualFeasibilitiesSSE += o.dualFeasibilitiesSSE;
    dynamicsViolationSSE += o.dynamicsViolationSSE;
    equalityConstraintsSSE += o.equalityConstraintsSSE;
    inequalityConstraintsSSE += o.inequalityConstraintsSSE;
    equalityLagrangian += o.equalityLagrangian;
    inequalityLagrangian += o.inequalityLagrangian;
    return *this;
  }
};

enum class Event : int { None = 0, PreEvent = 1, PostEvent = 2 };
struct AnnotatedTime {
  double time;
  Event event;
};

constexpr double kLimitEps = 1e-6;  // numeric_traits::limitEpsilon (ocs2_core/include/ocs2_core/NumericTraits.h:41)
constexpr double kWeakEps = 1e-9;   // numeric_traits::weakEpsilon  (:51)

inline double getIntervalStart(const AnnotatedTime& s) { return s.time + (s.event == Event::PostEvent ? kWeakEps : 0.0); }
inline double getIntervalEnd(const AnnotatedTime& e) { return e.time - (e.event == Event::PreEvent ? kWeakEps : 0.0); }
inline double getIntervalDuration(const AnnotatedTime& s, const AnnotatedTime& e) { return getIntervalEnd(e) - getIntervalStart(s); }

// lookup::findIndexInTimeArray (ocs2_core/include/ocs2_core/misc/Lookup.h:89-92)
inline int findIndexInTimeArray(const std::vector<double>& t, double time) {
  return static_cast<int>(std::lower_bound(t.begin(), t.end(), time) - t.begin());
}

inline std::vector<AnnotatedTime> timeDiscretizationWithEvents(double initTime, double finalTime, double dt,
                                                               const std::vector<double>& eventTimes,
                                                               double dt_min = 10.0 * kLimitEps) {
  std::vector<AnnotatedTime> td;
  td.push_back({initTime, Event::None});
  size_t nextEventIdx = static_cast<size_t>(findIndexInTimeArray(eventTimes, initTime));
  AnnotatedTime next = td.back();
  while (td.back().time < finalTime) {
    next.time = next.time + dt;
    next.event = Event::None;
    if (nextEventIdx < eventTimes.size() && next.time >= eventTimes[nextEventIdx]) {
      next.time = eventTimes[nextEventIdx];
      next.event = Event::PreEvent;
      nextEventIdx++;
    }
    if (next.time >= finalTime) {
      next.time = finalTime;
      next.event = Event::None;
    }
    if (next.time > td.back().time + dt_min)
      td.push_back(next);
    else
      td.back() = next;
  }
  if (td.front().event == Event::PreEvent) td.front().event = Event::PostEvent;
  std::vector<AnnotatedTime> out;
  out.reserve(2 * td.size());
  for (const auto& t : td) {
    out.push_back(t);
    if (t.event == Event::PreEvent) out.push_back({t.time, Event::PostEvent});
  }
  return out;
}

// ------------------------------------------------------------------------------------------------------
// Problem interface: what the solver asks of an OptimalControlProblem (ocs2_oc/.../OptimalControlProblem.h:48-138),
// indexed by shooting node k (reference-dependent data is per node; see wb_problem.hpp).
// ------------------------------------------------------------------------------------------------------
struct Ocp {
  int nx = 0, nu = 0;
  virtual ~Ocp() = default;
  virtual Vec flowMap(int k, double t, const Vec& x, const Vec& u) = 0;
  virtual LinApprox flowMapLin(int k, double t, const Vec& x, const Vec& u) = 0;
  virtual double cost(int k, double t, const Vec& x, const Vec& u) = 0;            // cost + soft constraints (intermediate)
  virtual QuadApprox costQuad(int k, double t, const Vec& x, const Vec& u) = 0;
  virtual double finalCost(int k, double t, const Vec& x) = 0;
  virtual QuadApprox finalCostQuad(int k, double t, const Vec& x) = 0;
  virtual Vec eqConstraint(int k, double t, const Vec& x, const Vec& u) { return {}; }
  virtual LinApprox eqConstraintLin(int k, double t, const Vec& x, const Vec& u) { return {}; }
  // event (pre-jump) node: identity jump map and no cost unless overridden (the humanoid OCPs define neither)
  virtual Vec jumpMap(int k, double t, const Vec& x) { return x; }
  virtual Mat jumpMapDx(int k, double t, const Vec& x) { return Mat::identity(nx); }
  virtual double eventCost(int k, double t, const Vec& x) { return 0.0; }
  virtual QuadApprox eventCostQuad(int k, double t, const Vec& x) {
    QuadApprox c;
    c.dfdxx = Mat(nx, nx);
    c.dfdx = vzero(nx);
    return c;
  }
};

// ---- integration ---------------------------------------------------------------------------------------
inline Vec rk4Discretization(Ocp& sys, int k, double t, const Vec& x, const Vec& u, double dt) {
  const double h2 = dt / 2.0, h6 = dt / 6.0, h3 = dt / 3.0;
  const Vec k1 = sys.flowMap(k, t, x, u);
  Vec tmp = x + h2 * k1;
  const Vec k2 = sys.flowMap(k, t + h2, tmp, u);
  tmp = x + h2 * k2;
  const Vec k3 = sys.flowMap(k, t + h2, tmp, u);
  tmp = x + dt * k3;
  const Vec k4 = sys.flowMap(k, t + dt, tmp, u);
  tmp = x;
  axpy(h6, k1, tmp);
  axpy(h3, k2, tmp);
  axpy(h3, k3, tmp);
  axpy(h6, k4, tmp);
  return tmp;
}

inline LinApprox rk4SensitivityDiscretization(Ocp& sys, int k, double t, const Vec& x, const Vec& u, double dt) {
  const double h2 = dt / 2.0, h6 = dt / 6.0, h3 = dt / 3.0;
  LinApprox k1 = sys.flowMapLin(k, t, x, u);
  Vec tmp = x + h2 * k1.f;
  LinApprox k2 = sys.flowMapLin(k, t + h2, tmp, u);
  tmp = x + h2 * k2.f;
  LinApprox k3 = sys.flowMapLin(k, t + h2, tmp, u);
  tmp = x + dt * k3.f;
  LinApprox k4 = sys.flowMapLin(k, t + dt, tmp, u);
  addTo(k2.dfdu, mul(k2.dfdx, k1.dfdu), h2);
  addTo(k3.dfdu, mul(k3.dfdx, k2.dfdu), h2);
  addTo(k4.dfdu, mul(k4.dfdx, k3.dfdu), dt);
  addTo(k2.dfdx, mul(k2.dfdx, k1.dfdx), h2);
  addTo(k3.dfdx, mul(k3.dfdx, k2.dfdx), h2);
  addTo(k4.dfdx, mul(k4.dfdx, k3.dfdx), dt);
  LinApprox out;
  out.dfdx = h6 * k1.dfdx + h3 * k2.dfdx + h3 * k3.dfdx + h6 * k4.dfdx;
  for (int i = 0; i < out.dfdx.r; ++i) out.dfdx(i, i) += 1.0;
  out.dfdu = h6 * k1.dfdu + h3 * k2.dfdu + h3 * k3.dfdu + h6 * k4.dfdu;
  out.f = x;
  axpy(h6, k1.f, out.f);
  axpy(h3, k2.f, out.f);
  axpy(h3, k3.f, out.f);
  axpy(h6, k4.f, out.f);
  return out;
}

// ---- projection ------------------------------------------------------------------------------------------
inline void changeOfInputVariables(QuadApprox& q, const Mat& Pu, const Mat& Px, const Vec& u0) {
  Mat P_plus_R_Px = q.dfdux + mul(q.dfduu, Px);
  Vec r_plus_R_u0 = q.dfdu + mul(q.dfduu, u0);
  addTo(q.dfdxx, mul(q.dfdux, Px, true, false));
  addTo(q.dfdxx, mul(Px, P_plus_R_Px, true, false));
  q.dfdx = q.dfdx + mul(q.dfdux, u0, true);
  q.dfdx = q.dfdx + mul(Px, r_plus_R_u0, true);
  q.f += 0.5 * dot(u0, r_plus_R_u0 + q.dfdu);
  q.dfdux = mul(Pu, P_plus_R_Px, true, false);
  Mat R_Pu = mul(q.dfduu, Pu);
  q.dfduu = mul(Pu, R_Pu, true, false);
  q.dfdu = mul(Pu, r_plus_R_u0, true);
}
inline void changeOfInputVariables(LinApprox& l, const Mat& Pu, const Mat& Px, const Vec& u0) {
  addTo(l.dfdx, mul(l.dfdu, Px));
  l.f = l.f + mul(l.dfdu, u0);
  l.dfdu = mul(l.dfdu, Pu);
}
// returns projection {dfdu = Pu, dfdx = Px, f = u0}
inline LinApprox luConstraintProjection(const LinApprox& constraint) {
  FullPivLU lu(constraint.dfdu);
  LinApprox p;
  p.dfdu = lu.kernel();
  p.dfdx = -1.0 * lu.solve(constraint.dfdx);
  p.f = -1.0 * lu.solve(constraint.f);
  return p;
}

// ---- QP: Riccati recursion as run by HPIPM through HpipmInterface ---------------------------------------
struct RiccatiSolution {
  std::vector<Vec> dx, du;     // dx[0] = delta_x0
  std::vector<Mat> P, K;       // cost-to-go Hessians (N+1), feedback gains (N)  [getRiccatiCostToGo / getRiccatiFeedback]
  std::vector<Vec> p, kff;     // cost-to-go gradients (N+1), feedforward (N)
  bool ok = true;
};

inline RiccatiSolution solveRiccati(const Vec& dx0, const std::vector<LinApprox>& dyn, const std::vector<QuadApprox>& cost,
                                    double reg_prim = 1e-12) {
  const int N = static_cast<int>(dyn.size());
  RiccatiSolution s;
  s.P.resize(N + 1);
  s.p.resize(N + 1);
  s.K.resize(N);
  s.kff.resize(N);
  std::vector<Mat> Lr(N);
  // terminal stage: HPIPM factorises Q_N + reg (P_N = Q_N + reg*I)
  s.P[N] = cost[N].dfdxx;
  for (int i = 0; i < s.P[N].r; ++i) s.P[N](i, i) += reg_prim;
  s.p[N] = cost[N].dfdx;
  for (int k = N - 1; k >= 0; --k) {
    const Mat& A = dyn[k].dfdx;
    const Mat& B = dyn[k].dfdu;
    const Vec& b = dyn[k].f;
    const int nu = B.c, nx = A.c;
    const Mat& Pn = s.P[k + 1];
    const Mat PA = mul(Pn, A);
    Vec Pb_p = mul(Pn, b) + s.p[k + 1];
    Mat Qt = cost[k].dfdxx + mul(A, PA, true, false);
    for (int i = 0; i < nx; ++i) Qt(i, i) += reg_prim;
    Vec qt = cost[k].dfdx + mul(A, Pb_p, true);
    if (nu == 0) {  // event node: no input (Transcription.cpp:156-192)
      s.P[k] = Qt;
      s.p[k] = qt;
      s.K[k] = Mat(0, nx);
      s.kff[k] = Vec();
      continue;
    }
    const Mat PB = mul(Pn, B);
    Mat Rt = cost[k].dfduu + mul(B, PB, true, false);
    for (int i = 0; i < nu; ++i) Rt(i, i) += reg_prim;
    Mat St = cost[k].dfdux + mul(B, PA, true, false);  // nu x nx
    Vec rt = cost[k].dfdu + mul(B, Pb_p, true);
    Lr[k] = Rt;
    if (!choleskyLower(Lr[k])) {
      s.ok = false;
      return s;
    }
    Mat LiS = St;  // Lr^-1 S~   (Ls' in HPIPM's notation)
    solveLower(Lr[k], LiS);
    Mat Lir = asCol(rt);
    solveLower(Lr[k], Lir);
    s.P[k] = Qt - mul(LiS, LiS, true, false);
    // keep P exactly symmetric, as the packed Cholesky factor product in HPIPM is
    for (int i = 0; i < nx; ++i)
      for (int j = i + 1; j < nx; ++j) {
        const double m = 0.5 * (s.P[k](i, j) + s.P[k](j, i));
        s.P[k](i, j) = s.P[k](j, i) = m;
      }
    s.p[k] = qt - mul(LiS, Lir.a, true);
    Mat Kk = LiS;
    solveLowerT(Lr[k], Kk);
    s.K[k] = -1.0 * Kk;
    Mat kk = Lir;
    solveLowerT(Lr[k], kk);
    s.kff[k] = -1.0 * kk.a;
  }
  // forward pass
  s.dx.resize(N + 1);
  s.du.resize(N);
  s.dx[0] = dx0;
  for (int k = 0; k < N; ++k) {
    const int nu = dyn[k].dfdu.c;
    if (nu > 0) {
      s.du[k] = mul(s.K[k], s.dx[k]) + s.kff[k];
      s.dx[k + 1] = mul(dyn[k].dfdx, s.dx[k]) + mul(dyn[k].dfdu, s.du[k]) + dyn[k].f;
    } else {
      s.du[k] = Vec();
      s.dx[k + 1] = mul(dyn[k].dfdx, s.dx[k]) + dyn[k].f;
    }
  }
  return s;
}

// ---- filter line search -------------------------------------------------------------------------------------
enum class StepType : int { UNKNOWN = 0, CONSTRAINT = 1, DUAL = 2, COST = 3, ZERO = 4 };
struct FilterLinesearch {
  double g_max = 1e6, g_min = 1e-6, gamma_c = 1e-6, armijoFactor = 1e-4;
  static double totalConstraintViolation(const PerformanceIndex& p) { return std::sqrt(p.dynamicsViolationSSE + p.equalityConstraintsSSE); }
  std::pair<bool, StepType> acceptStep(const PerformanceIndex& base, const PerformanceIndex& step, double armijoDescentMetric) const {
    const double g0 = totalConstraintViolation(base), g1 = totalConstraintViolation(step);
    if (g1 > g_max) return {g1 < (1.0 - gamma_c) * g0, StepType::CONSTRAINT};
    if (g1 < g_min && g0 < g_min && armijoDescentMetric < 0.0) return {step.merit < base.merit + armijoFactor * armijoDescentMetric, StepType::COST};
    return {step.merit < base.merit - gamma_c * g0 || g1 < (1.0 - gamma_c) * g0, StepType::DUAL};
  }
};

struct SqpSettings {  // sqp::Settings defaults (ocs2_sqp/ocs2_sqp/include/ocs2_sqp/SqpSettings.h:40-87)
  int sqpIteration = 10;
  double deltaTol = 1e-6, costTol = 1e-4;
  double alpha_decay = 0.5, alpha_min = 1e-4, gamma_c = 1e-6, g_max = 1e6, g_min = 1e-6, armijoFactor = 1e-4;
  double dt = 0.01;
  bool projectStateInputEqualityConstraints = true;
  bool useFeedbackPolicy = true;
  double reg_prim = 1e-12;  // hpipm_catkin/include/hpipm_catkin/HpipmInterfaceSettings.h:45-56
};

enum class Convergence : int { FALSE_ = 0, ITERATIONS = 1, STEPSIZE = 2, METRICS = 3, PRIMAL = 4 };

struct StepInfo {
  double stepSize = 0, dx_norm = 0, du_norm = 0, totalConstraintViolationAfterStep = 0;
  StepType stepType = StepType::UNKNOWN;
  PerformanceIndex performanceAfterStep;
};

struct IterationLog {
  PerformanceIndex baseline;
  StepInfo step;
  double armijoDescentMetric = 0;
  Convergence convergence = Convergence::FALSE_;
};

inline double trajectoryNorm(const std::vector<Vec>& v) {  // Helpers.h:43-50
  double s = 0;
  for (const auto& e : v) s += sqnorm(e);
  return std::sqrt(s);
}

class SqpOracle {
 public:
  SqpOracle(Ocp& ocp, SqpSettings s) : ocp_(ocp), settings_(s) {
    ls_.g_max = s.g_max;
    ls_.g_min = s.g_min;
    ls_.gamma_c = s.gamma_c;
    ls_.armijoFactor = s.armijoFactor;
  }

  // LQ data of the last iteration (after projection), exposed for block-level parity tests
  std::vector<LinApprox> dynamics, projection, constraintsRaw;
  std::vector<QuadApprox> cost;
  std::vector<LinApprox> dynamicsRaw;  // before projection
  std::vector<QuadApprox> costRaw;
  RiccatiSolution qp;
  std::vector<Vec> deltaX, deltaU;  // remapped
  std::vector<Mat> Kgain;           // remapped feedback gains (toPrimalSolution with useFeedbackPolicy)
  std::vector<IterationLog> log;

  PerformanceIndex setupQuadraticSubproblem(const std::vector<AnnotatedTime>& time, const Vec& initState, const std::vector<Vec>& x,
                                            const std::vector<Vec>& u, bool keepRaw = false) {
    const int N = static_cast<int>(time.size()) - 1;
    cost.assign(N + 1, {});
    dynamics.assign(N, {});
    projection.assign(N, {});
    if (keepRaw) {
      dynamicsRaw.assign(N, {});
      costRaw.assign(N + 1, {});
      constraintsRaw.assign(N, {});
    }
    PerformanceIndex perf;
    for (int i = 0; i < N; ++i) {
      if (time[i].event == Event::PreEvent) {
        // setupEventNode (Transcription.cpp:156-192)
        LinApprox d;
        d.dfdx = ocp_.jumpMapDx(i, time[i].time, x[i]);
        d.dfdu = Mat(ocp_.nx, 0);
        d.f = ocp_.jumpMap(i, time[i].time, x[i]) - x[i + 1];
        QuadApprox c = ocp_.eventCostQuad(i, time[i].time, x[i]);
        c.dfdux = Mat(0, ocp_.nx);
        c.dfduu = Mat(0, 0);
        c.dfdu = Vec();
        PerformanceIndex p;
        p.cost = c.f;
        p.dynamicsViolationSSE = sqnorm(d.f);
        perf += p;
        if (keepRaw) {
          dynamicsRaw[i] = d;
          costRaw[i] = c;
        }
        dynamics[i] = std::move(d);
        cost[i] = std::move(c);
      } else {
        const double ti = getIntervalStart(time[i]);
        const double dt = getIntervalDuration(time[i], time[i + 1]);
        LinApprox d = rk4SensitivityDiscretization(ocp_, i, ti, x[i], u[i], dt);
        d.f = d.f - x[i + 1];
        QuadApprox c = ocp_.costQuad(i, ti, x[i], u[i]);
        c.dfdxx = dt * c.dfdxx;
        c.dfdux = dt * c.dfdux;
        c.dfduu = dt * c.dfduu;
        c.dfdx = dt * c.dfdx;
        c.dfdu = dt * c.dfdu;
        c.f *= dt;
        LinApprox g = ocp_.eqConstraintLin(i, ti, x[i], u[i]);
        PerformanceIndex p;
        p.dynamicsViolationSSE = dt * sqnorm(d.f);
        p.cost = c.f;
        p.equalityConstraintsSSE = dt * sqnorm(g.f);
        perf += p;
        if (keepRaw) {
          dynamicsRaw[i] = d;
          costRaw[i] = c;
          constraintsRaw[i] = g;
        }
        if (settings_.projectStateInputEqualityConstraints && !g.f.empty()) {
          projection[i] = luConstraintProjection(g);
          changeOfInputVariables(d, projection[i].dfdu, projection[i].dfdx, projection[i].f);
          changeOfInputVariables(c, projection[i].dfdu, projection[i].dfdx, projection[i].f);
        } else if (!g.f.empty()) {
          throw std::runtime_error("oracle: constraints-in-QP path not restated (G1 uses projection)");
        }
        dynamics[i] = std::move(d);
        cost[i] = std::move(c);
      }
    }
    {
      const double tN = getIntervalStart(time[N]);
      QuadApprox c = ocp_.finalCostQuad(N, tN, x[N]);
      PerformanceIndex p;
      p.cost = c.f;
      perf += p;
      if (keepRaw) costRaw[N] = c;
      cost[N] = std::move(c);
    }
    perf.dynamicsViolationSSE += sqnorm(initState - x.front());
    perf.merit = perf.cost + perf.equalityLagrangian + perf.inequalityLagrangian;
    return perf;
  }

  PerformanceIndex computePerformance(const std::vector<AnnotatedTime>& time, const Vec& initState, const std::vector<Vec>& x,
                                      const std::vector<Vec>& u) {
    const int N = static_cast<int>(time.size()) - 1;
    PerformanceIndex perf;
    for (int i = 0; i < N; ++i) {
      PerformanceIndex p;
      if (time[i].event == Event::PreEvent) {
        p.cost = ocp_.eventCost(i, time[i].time, x[i]);
        p.dynamicsViolationSSE = sqnorm(ocp_.jumpMap(i, time[i].time, x[i]) - x[i + 1]);
      } else {
        const double ti = getIntervalStart(time[i]);
        const double dt = getIntervalDuration(time[i], time[i + 1]);
        Vec dv = rk4Discretization(ocp_, i, ti, x[i], u[i], dt) - x[i + 1];
        p.cost = dt * ocp_.cost(i, ti, x[i], u[i]);
        p.dynamicsViolationSSE = dt * sqnorm(dv);
        p.equalityConstraintsSSE = dt * sqnorm(ocp_.eqConstraint(i, ti, x[i], u[i]));
      }
      perf += p;
    }
    PerformanceIndex pN;
    pN.cost = ocp_.finalCost(N, getIntervalStart(time[N]), x[N]);
    perf += pN;
    perf.dynamicsViolationSSE += sqnorm(initState - x.front());
    perf.merit = perf.cost + perf.equalityLagrangian + perf.inequalityLagrangian;
    return perf;
  }

  // one full runImpl on a given discretisation and initial guess; x,u updated in place
  Convergence run(const std::vector<AnnotatedTime>& time, const Vec& initState, std::vector<Vec>& x, std::vector<Vec>& u,
                  bool keepRaw = false) {
    const int N = static_cast<int>(time.size()) - 1;
    log.clear();
    int iter = 0;
    Convergence conv = Convergence::FALSE_;
    while (conv == Convergence::FALSE_) {
      IterationLog entry;
      const PerformanceIndex baseline = setupQuadraticSubproblem(time, initState, x, u, keepRaw);
      entry.baseline = baseline;
      const Vec dx0 = initState - x[0];
      qp = solveRiccati(dx0, dynamics, cost, settings_.reg_prim);
      if (!qp.ok) throw std::runtime_error("[SqpOracle] Failed to solve QP");
      // armijoDescentMetric on the projected QP (FilterLinesearch.cpp:76-89)
      double armijo = 0.0;
      for (int i = 0; i <= N; ++i) {
        armijo += dot(cost[i].dfdx, qp.dx[i]);
        if (i < N && !cost[i].dfdu.empty()) armijo += dot(cost[i].dfdu, qp.du[i]);
      }
      entry.armijoDescentMetric = armijo;
      deltaX = qp.dx;
      deltaU = qp.du;
      Kgain = qp.K;
      for (int i = 0; i < N; ++i) {
        if (!projection[i].f.empty()) {  // remapProjectedInput / remapProjectedGain
          deltaU[i] = mul(projection[i].dfdu, qp.du[i]) + projection[i].f + mul(projection[i].dfdx, deltaX[i]);
          Kgain[i] = mul(projection[i].dfdu, qp.K[i]) + projection[i].dfdx;
        }
      }
      // takeStep
      const double dUn = trajectoryNorm(deltaU), dXn = trajectoryNorm(deltaX);
      double alpha = 1.0;
      StepInfo info;
      bool accepted = false;
      do {
        std::vector<Vec> xn(x.size()), un(u.size());
        for (size_t i = 0; i < x.size(); ++i) {
          xn[i] = x[i];
          axpy(alpha, deltaX[i], xn[i]);
        }
        for (size_t i = 0; i < u.size(); ++i) {
          un[i] = u[i];
          if (!deltaU[i].empty()) axpy(alpha, deltaU[i], un[i]);
        }
        const PerformanceIndex pn = computePerformance(time, initState, xn, un);
        auto res = ls_.acceptStep(baseline, pn, alpha * armijo);
        if (res.first) {
          x = std::move(xn);
          u = std::move(un);
          info.stepSize = alpha;
          info.stepType = res.second;
          info.dx_norm = alpha * dXn;
          info.du_norm = alpha * dUn;
          info.performanceAfterStep = pn;
          info.totalConstraintViolationAfterStep = FilterLinesearch::totalConstraintViolation(pn);
          accepted = true;
          break;
        }
        alpha *= settings_.alpha_decay;
        if (alpha * dXn < settings_.deltaTol && alpha * dUn < settings_.deltaTol) break;
      } while (alpha >= settings_.alpha_min);
      if (!accepted) {
        info.stepSize = 0.0;
        info.stepType = StepType::ZERO;
        info.performanceAfterStep = baseline;
        info.totalConstraintViolationAfterStep = FilterLinesearch::totalConstraintViolation(baseline);
      }
      // checkConvergence
      if (iter + 1 >= settings_.sqpIteration)
        conv = Convergence::ITERATIONS;
      else if (info.stepSize < settings_.alpha_min)
        conv = Convergence::STEPSIZE;
      else if (std::fabs(info.performanceAfterStep.merit - baseline.merit) < settings_.costTol &&
               FilterLinesearch::totalConstraintViolation(info.performanceAfterStep) < settings_.g_min)
        conv = Convergence::METRICS;
      else if (info.dx_norm < settings_.deltaTol && info.du_norm < settings_.deltaTol)
        conv = Convergence::PRIMAL;
      entry.step = info;
      entry.convergence = conv;
      log.push_back(entry);
      ++iter;
    }
    return conv;
  }

  const SqpSettings& settings() const { return settings_; }

 private:
  Ocp& ocp_;
  SqpSettings settings_;
  FilterLinesearch ls_;
};

}  // namespace orc
