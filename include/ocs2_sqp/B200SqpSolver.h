/*
 * ocs2::B200SqpSolver -- drop-in for ocs2::SqpSolver (lib/ocs2_ros2/ocs2_sqp/ocs2_sqp/include/ocs2_sqp/SqpSolver.h:60-103) for the humanoid
 * whole-body / centroidal OCP family: a SolverBase (ocs2_oc/include/ocs2_oc/oc_solver/SolverBase.h:52-273) whose runImpl hands the SQP
 * iteration to libb200sqp.so through the C ABI (include/b200sqp.h).  Header-only; it is written against the ocs2 headers a workspace of
 * manumerous/wb_humanoid_mpc already has.  In this repository it is compiled and run against the stand-in headers of tests/stubs/
 * (tests/test_shim.py; the image has no Eigen / Boost / Pinocchio / ROS 2), using only the part of the Eigen API those stand-ins also offer
 * (size(), data(), operator[], vector_t(n), matrix_t(r, c), operator()(i, j)).
 *
 * What stays on the host, unchanged from the reference: SolverBase::preRun (reference manager: gait -> mode schedule, swing planner, target
 * trajectories), timeDiscretizationWithEvents, initializeStateInputTrajectories (warm start + Initializer), toPrimalSolution.  What moves to
 * the GPU: everything between SqpSolver.cpp:219 and :271 (LQ approximation, QP, line search, convergence).
 *
 * The OptimalControlProblem is a bag of virtual term objects backed by CppAD-generated code and cannot be shipped to a GPU; the device
 * constants come from a b200sqp_model_desc derived from the same URDF + task.info the interface was built from
 * (wb_humanoid_mpc_b200/host/model_file.hpp reads the flat file tools/make_model_data.py writes; SURVEY.md section 8b "the catch").
 */
#pragma once

#include <b200sqp.h>

#include <cmath>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <humanoid_common_mpc/reference_manager/SwitchedModelReferenceManager.h>
#include <ocs2_core/Types.h>
#include <ocs2_core/control/FeedforwardController.h>
#include <ocs2_core/control/LinearController.h>
#include <ocs2_core/initialization/Initializer.h>
#include <ocs2_oc/multiple_shooting/Initialization.h>
#include <ocs2_oc/oc_data/PerformanceIndex.h>
#include <ocs2_oc/oc_data/PrimalSolution.h>
#include <ocs2_oc/oc_data/TimeDiscretization.h>
#include <ocs2_oc/oc_problem/OptimalControlProblem.h>
#include <ocs2_oc/oc_solver/SolverBase.h>
#include <ocs2_sqp/SqpSettings.h>

namespace ocs2 {
namespace b200 {

/** sqp::Settings -> b200sqp_settings (the fields the hot path reads, SqpSettings.h:40-87) */
inline b200sqp_settings toB200Settings(const sqp::Settings& s) {
  b200sqp_settings st;
  b200sqp_default_settings(&st);
  st.sqp_iteration = static_cast<int32_t>(s.sqpIteration);
  st.delta_tol = s.deltaTol;
  st.cost_tol = s.costTol;
  st.alpha_decay = s.alpha_decay;
  st.alpha_min = s.alpha_min;
  st.gamma_c = s.gamma_c;
  st.g_max = s.g_max;
  st.g_min = s.g_min;
  st.armijo_factor = s.armijoFactor;
  st.reg_prim = s.hpipmSettings.reg_prim;
  st.use_feedback_policy = s.useFeedbackPolicy ? 1 : 0;
  st.create_value_function = s.createValueFunction ? 1 : 0;
  st.global_step = 0;
  return st;
}

/** flat per-node arrays of one instance in the layout of b200sqp_upload_instances */
struct NodeArrays {
  std::vector<double> t, x, u, swing, impact, armPhase, xref, K;
  std::vector<uint8_t> event, contact;
};

/**
 * What setupQuadraticSubproblem's terms read from the reference manager at every node (SwitchedModelReferenceManager.cpp:110-154,
 * WBMpcPreComputation.cpp:68-113), evaluated once per node on the host:
 * contact flags, swing-z reference (position, velocity, acceleration), impact proximity factor, arm-swing phase, target state.
 */
inline NodeArrays fillNodeArrays(const std::vector<AnnotatedTime>& time, const humanoid::SwitchedModelReferenceManager& rm, const vector_array_t& x,
                                 const vector_array_t& u, size_t nx, size_t nu) {
  const size_t n = time.size();
  NodeArrays a;
  a.t.resize(n);
  a.event.resize(n);
  a.contact.resize(2 * n);
  a.swing.resize(6 * n);
  a.impact.resize(2 * n);
  a.armPhase.resize(n);
  a.xref.resize(n * nx);
  a.x.resize(n * nx);
  a.u.assign((n - 1) * nu, 0.0);
  const auto& planner = *rm.getSwingTrajectoryPlanner();
  const TargetTrajectories& targets = rm.getTargetTrajectories();
  constexpr double kPi = 3.14159265358979323846;
  for (size_t i = 0; i < n; ++i) {
    a.t[i] = time[i].time;
    a.event[i] = time[i].event == AnnotatedTime::Event::PreEvent ? 1 : (time[i].event == AnnotatedTime::Event::PostEvent ? 2 : 0);
    const scalar_t ti = getIntervalStart(time[i]);
    const auto c = rm.getContactFlags(ti);
    for (size_t leg = 0; leg < 2; ++leg) {
      a.contact[2 * i + leg] = c[leg] ? 1 : 0;
      a.swing[(2 * i + leg) * 3 + 0] = planner.getZpositionConstraint(leg, ti);
      a.swing[(2 * i + leg) * 3 + 1] = planner.getZvelocityConstraint(leg, ti);
      a.swing[(2 * i + leg) * 3 + 2] = planner.getZaccelerationConstraint(leg, ti);
      a.impact[2 * i + leg] = planner.getImpactProximityFactor(leg, ti);
    }
    a.armPhase[i] = std::sin(2.0 * kPi * (rm.getPhaseVariable(ti) - 0.15));   // SwitchedModelReferenceManager.cpp:110-135
    const vector_t xr = targets.getDesiredState(ti);
    for (size_t k = 0; k < nx; ++k) {
      a.xref[i * nx + k] = xr[k];
      a.x[i * nx + k] = x[i][k];
    }
    if (i + 1 < n && u[i].size() == nu)   // (no input at a pre-event node)
      for (size_t k = 0; k < nu; ++k) a.u[i * nu + k] = u[i][k];
  }
  return a;
}

/** multiple_shooting::toPrimalSolution (ocs2_oc/src/multiple_shooting/Helpers.cpp:60-120): the pre-event node's input is the previous one */
inline PrimalSolution toPrimalSolution(const std::vector<AnnotatedTime>& time, const ModeSchedule& modeSchedule, const NodeArrays& a, size_t nx, size_t nu,
                                       bool feedback) {
  const size_t n = time.size();
  PrimalSolution p;
  p.modeSchedule_ = modeSchedule;
  vector_array_t uff;
  matrix_array_t gains;
  for (size_t i = 0; i < n; ++i) {
    p.timeTrajectory_.push_back(time[i].time);
    vector_t xi(nx);
    for (size_t k = 0; k < nx; ++k) xi[k] = a.x[i * nx + k];
    p.stateTrajectory_.push_back(xi);
    // the terminal node repeats the last input; a pre-event node repeats the input before it
    size_t src = (i + 1 < n) ? i : n - 2;
    if (time[i].event == AnnotatedTime::Event::PreEvent && i > 0) src = i - 1;
    vector_t ui(nu);
    for (size_t k = 0; k < nu; ++k) ui[k] = a.u[src * nu + k];
    p.inputTrajectory_.push_back(ui);
    if (time[i].event == AnnotatedTime::Event::PreEvent) p.postEventIndices_.push_back(i + 1);
    if (feedback) {
      matrix_t K(nu, nx);
      vector_t bias(nu);
      for (size_t c = 0; c < nx; ++c)
        for (size_t r = 0; r < nu; ++r) K(r, c) = a.K[src * nu * nx + r + nu * c];
      for (size_t r = 0; r < nu; ++r) {
        scalar_t s = ui[r];
        for (size_t c = 0; c < nx; ++c) s -= K(r, c) * xi[c];   // uff = u - K x (SqpSolver.cpp:338-340)
        bias[r] = s;
      }
      gains.push_back(K);
      uff.push_back(bias);
    }
  }
  if (feedback) p.controllerPtr_.reset(new LinearController(p.timeTrajectory_, uff, gains));
  else p.controllerPtr_.reset(new FeedforwardController(p.timeTrajectory_, p.inputTrajectory_));
  return p;
}

/** b200sqp_iter_log -> the PerformanceIndex the reference logs after the step of that iteration (SqpSolver.cpp:251-268) */
inline PerformanceIndex toPerformanceIndex(const b200sqp_iter_log& l) {
  PerformanceIndex p;
  p.merit = l.merit;
  p.cost = l.cost;
  p.dynamicsViolationSSE = l.dyn_sse;
  p.equalityConstraintsSSE = l.eq_sse;
  return p;
}

}  // namespace b200

class B200SqpSolver : public SolverBase {
 public:
  /**
   * Mirrors SqpSolver(settings, optimalControlProblem, initializer) (SqpSolver.cpp:58-83) plus the device constants of the OCP family.
   * centroidal = nullptr: whole-body OCP (WBMpcInterface.cpp:131-199); otherwise the centroidal OCP (CentroidalMpcInterface.cpp:150-237).
   * Throws std::runtime_error when the device or the library is unavailable (there is no CPU fallback).
   */
  B200SqpSolver(sqp::Settings settings, const OptimalControlProblem& optimalControlProblem, const Initializer& initializer, const b200sqp_model_desc& model,
                const b200sqp_cen_desc* centroidal = nullptr, int device = 0)
      : settings_(std::move(settings)), ocp_(optimalControlProblem), initializerPtr_(initializer.clone()), nx_(centroidal ? 12 + model.nj : 2 * (6 + model.nj)),
        nu_(12 + model.nj) {
    const b200sqp_settings st = b200::toB200Settings(settings_);
    if (centroidal) check(b200sqp_cen_create(&model, centroidal, &st, device, &handle_));
    else check(b200sqp_create(&model, &st, device, &handle_));
  }
  ~B200SqpSolver() override { b200sqp_destroy(handle_); }
  B200SqpSolver(const B200SqpSolver&) = delete;
  B200SqpSolver& operator=(const B200SqpSolver&) = delete;

  void reset() override {   // SqpSolver::reset (SqpSolver.cpp:112-127)
    primalSolution_ = PrimalSolution();
    performanceIndeces_.clear();
    numIterations_ = 0;
  }

  scalar_t getFinalTime() const override { return primalSolution_.timeTrajectory_.back(); }
  void getPrimalSolution(scalar_t, PrimalSolution* primalSolutionPtr) const override { *primalSolutionPtr = primalSolution_; }
  const ProblemMetrics& getSolutionMetrics() const override { return problemMetrics_; }
  size_t getNumIterations() const override { return numIterations_; }
  const OptimalControlProblem& getOptimalControlProblem() const override { return ocp_; }
  const PerformanceIndex& getPerformanceIndeces() const override { return getIterationsLog().back(); }
  const std::vector<PerformanceIndex>& getIterationsLog() const override {
    if (performanceIndeces_.empty()) throw std::runtime_error("[B200SqpSolver]: No performance log yet, no problem solved yet?");   // SqpSolver.cpp:164-170
    return performanceIndeces_;
  }

  /** SqpSolver::getValueFunction (SqpSolver.cpp:172-191): quadratic model around the nearest node at or before `time` */
  ScalarFunctionQuadraticApproximation getValueFunction(scalar_t time, const vector_t& state) const override {
    if (valueP_.empty()) throw std::runtime_error("[B200SqpSolver] Value function is empty! Is createValueFunction true and did the solver run?");
    const auto& t = primalSolution_.timeTrajectory_;
    size_t i = 0;
    while (i + 1 < t.size() && t[i + 1] <= time) ++i;
    ScalarFunctionQuadraticApproximation v;
    v.dfdxx = matrix_t(nx_, nx_);
    v.dfdx = vector_t(nx_);
    for (size_t c = 0; c < nx_; ++c)
      for (size_t r = 0; r < nx_; ++r) v.dfdxx(r, c) = valueP_[(i * nx_ + c) * nx_ + r];
    for (size_t r = 0; r < nx_; ++r) {
      scalar_t s = valueP_.empty() ? 0.0 : valuep_[i * nx_ + r];
      for (size_t c = 0; c < nx_; ++c) s += v.dfdxx(r, c) * state[c];   // dfdx = p (re-centred) + P x
      v.dfdx[r] = s;
    }
    return v;
  }
  ScalarFunctionQuadraticApproximation getHamiltonian(scalar_t, const vector_t&, const vector_t&) override {
    throw std::runtime_error("[B200SqpSolver] getHamiltonian() not available yet.");   // SqpSolver.h:82-84
  }
  vector_t getStateInputEqualityConstraintLagrangian(scalar_t, const vector_t&) const override {
    throw std::runtime_error("[B200SqpSolver] getStateInputEqualityConstraintLagrangian() not available yet.");   // SqpSolver.h:86-88
  }
  MultiplierCollection getIntermediateDualSolution(scalar_t) const override {
    throw std::runtime_error("[B200SqpSolver] getIntermediateDualSolution() not available yet.");   // SqpSolver.h:90-92
  }

  /** SqpSolver::getBenchmarks (SqpSolver.h:97-102): device ms of {LQ approximation, solve QP, line search, projection share of LQ} of the last run */
  std::array<float, 4> getBenchmarks() const {
    std::array<float, 4> ms{};
    check(b200sqp_get_stage_times(handle_, ms.data()));
    return ms;
  }
  std::string getBenchmarkingInfo() const override {
    const auto ms = getBenchmarks();
    return "[B200SqpSolver] device ms: LQ approximation " + std::to_string(ms[0]) + ", solve QP " + std::to_string(ms[1]) + ", line search " + std::to_string(ms[2]);
  }

 private:
  void runImpl(scalar_t initTime, const vector_t& initState, scalar_t finalTime) override {
    auto* rm = dynamic_cast<const humanoid::SwitchedModelReferenceManager*>(&getReferenceManager());
    if (rm == nullptr) throw std::runtime_error("[B200SqpSolver] the reference manager must be a humanoid::SwitchedModelReferenceManager");
    // host work of the reference, unchanged (SqpSolver.cpp:201-217)
    const ModeSchedule& modeSchedule = rm->getModeSchedule();
    const auto time = timeDiscretizationWithEvents(initTime, finalTime, settings_.dt, modeSchedule.eventTimes);
    vector_array_t x, u;
    multiple_shooting::initializeStateInputTrajectories(initState, time, primalSolution_, *initializerPtr_, x, u);
    b200::NodeArrays nodes = b200::fillNodeArrays(time, *rm, x, u, nx_, nu_);
    const int n = static_cast<int>(time.size());
    if (n != nodesOnDevice_) {
      check(b200sqp_set_batch(handle_, 1, n));
      nodesOnDevice_ = n;
    }
    check(b200sqp_upload_instances(handle_, initState.data(), nodes.x.data(), nodes.u.data(), nodes.t.data(), nodes.event.data(), nodes.contact.data(),
                                   nodes.swing.data(), nodes.impact.data(), nodes.armPhase.data(), nodes.xref.data()));
    check(b200sqp_solve(handle_, nullptr));
    std::vector<b200sqp_iter_log> log(settings_.sqpIteration);
    int32_t nIter = 0, status = 0;
    if (settings_.useFeedbackPolicy) nodes.K.resize(static_cast<size_t>(n - 1) * nu_ * nx_);
    check(b200sqp_download(handle_, nodes.x.data(), nodes.u.data(), settings_.useFeedbackPolicy ? nodes.K.data() : nullptr, log.data(), &nIter, &status));
    if (settings_.createValueFunction) {
      valueP_.resize(static_cast<size_t>(n) * nx_ * nx_);
      valuep_.resize(static_cast<size_t>(n) * nx_);
      check(b200sqp_download_value_function(handle_, valueP_.data(), valuep_.data()));
    }
    primalSolution_ = b200::toPrimalSolution(time, modeSchedule, nodes, nx_, nu_, settings_.useFeedbackPolicy);
    performanceIndeces_.clear();
    for (int i = 0; i < nIter; ++i) performanceIndeces_.push_back(b200::toPerformanceIndex(log[i]));
    numIterations_ += static_cast<size_t>(nIter);
  }
  void runImpl(scalar_t initTime, const vector_t& initState, scalar_t finalTime, const ControllerBase* externalControllerPtr) override {
    if (externalControllerPtr == nullptr) runImpl(initTime, initState, finalTime);
    else throw std::runtime_error("[B200SqpSolver::run] This solver does not support external controller!");   // SqpSolver.h:109-115
  }
  void runImpl(scalar_t initTime, const vector_t& initState, scalar_t finalTime, const PrimalSolution& primalSolution) override {
    primalSolution_ = primalSolution;   // SqpSolver.h:117-121
    runImpl(initTime, initState, finalTime);
  }
  /** "[SqpSolver] Failed to solve QP" and every other failure surface as std::runtime_error (SqpSolver.cpp:306-308) */
  static void check(int rc) {
    if (rc != 0) throw std::runtime_error(std::string("[B200SqpSolver] ") + b200sqp_last_error());
  }

  sqp::Settings settings_;
  OptimalControlProblem ocp_;
  std::unique_ptr<Initializer> initializerPtr_;
  size_t nx_, nu_;
  b200sqp_handle handle_ = nullptr;
  int nodesOnDevice_ = 0;
  PrimalSolution primalSolution_;
  std::vector<PerformanceIndex> performanceIndeces_;
  ProblemMetrics problemMetrics_;
  std::vector<double> valueP_, valuep_;
  size_t numIterations_ = 0;
};

}  // namespace ocs2
