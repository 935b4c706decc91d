// TEST SCAFFOLDING: minimal stand-ins for the ocs2 / humanoid_common_mpc headers that include/ocs2_sqp/B200SqpSolver.h is written against, so
// that the shim compiles and runs in this image (no Eigen, Boost, Pinocchio, ROS 2).  Each declaration keeps the name, signature and meaning
// of the reference declaration it stands in for (cited); only what the shim touches exists.  The algorithmic pieces (time grid, warm start,
// gait schedule, swing planner) forward to this repository's C++ host layer, which is pinned on the oracle (tests/test_host_cpp.py,
// tests/test_oracle_spreading.py).  A workspace with the real ocs2 uses its own headers instead of this directory.
#pragma once
#include <array>
#include <cstddef>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../wb_humanoid_mpc_b200/host/references.hpp"

namespace ocs2 {

// ---- ocs2_core/Types.h:40-75 (Eigen typedefs there; here the subset of the Eigen API the shim uses) --------------------------------------
using scalar_t = double;
class vector_t {
 public:
  vector_t() = default;
  explicit vector_t(std::size_t n) : v_(n) {}
  vector_t(const std::vector<double>& v) : v_(v) {}
  static vector_t Zero(std::size_t n) {
    vector_t r(n);
    r.setZero();
    return r;
  }
  std::size_t size() const { return v_.size(); }
  double* data() { return v_.data(); }
  const double* data() const { return v_.data(); }
  double& operator[](std::size_t i) { return v_[i]; }
  double operator[](std::size_t i) const { return v_[i]; }
  double& operator()(std::size_t i) { return v_[i]; }
  double operator()(std::size_t i) const { return v_[i]; }
  void setZero() { std::fill(v_.begin(), v_.end(), 0.0); }
  void resize(std::size_t n) { v_.resize(n); }
  const std::vector<double>& std() const { return v_; }

 private:
  std::vector<double> v_;
};
class matrix_t {  // column-major, like Eigen's default
 public:
  matrix_t() = default;
  matrix_t(std::size_t r, std::size_t c) : r_(r), c_(c), v_(r * c) {}
  std::size_t rows() const { return r_; }
  std::size_t cols() const { return c_; }
  double* data() { return v_.data(); }
  const double* data() const { return v_.data(); }
  double& operator()(std::size_t i, std::size_t j) { return v_[i + j * r_]; }
  double operator()(std::size_t i, std::size_t j) const { return v_[i + j * r_]; }
  void setZero() { std::fill(v_.begin(), v_.end(), 0.0); }

 private:
  std::size_t r_ = 0, c_ = 0;
  std::vector<double> v_;
};
using scalar_array_t = std::vector<scalar_t>;
using size_array_t = std::vector<std::size_t>;
using vector_array_t = std::vector<vector_t>;
using matrix_array_t = std::vector<matrix_t>;
struct ScalarFunctionQuadraticApproximation {  // Types.h:145-157
  matrix_t dfdxx, dfdux, dfduu;
  vector_t dfdx, dfdu;
  scalar_t f = 0.0;
};
struct MultiplierCollection {};
struct ProblemMetrics {};
struct DualSolution {};
struct OptimalControlProblem {};  // ocs2_oc/oc_problem/OptimalControlProblem.h (a bag of term collections; opaque to the shim)

// ---- ocs2_core/reference/ModeSchedule.h:45-75, TargetTrajectories.h:41-70 ----------------------------------------------------------------
struct ModeSchedule {
  scalar_array_t eventTimes;
  size_array_t modeSequence{0};
  std::size_t modeAtTime(scalar_t t) const {
    std::size_t i = 0;
    while (i < eventTimes.size() && eventTimes[i] < t) ++i;  // lookup::findIndexInTimeArray (lower bound)
    return modeSequence[i];
  }
  void clear() {
    eventTimes.clear();
    modeSequence = {0};
  }
};
inline void swap(ModeSchedule& a, ModeSchedule& b) { std::swap(a, b); }
struct TargetTrajectories {
  scalar_array_t timeTrajectory;
  vector_array_t stateTrajectory, inputTrajectory;
  vector_t getDesiredState(scalar_t t) const {
    b200sqp::host::TargetTrajectories tt;
    tt.timeTrajectory = timeTrajectory;
    for (const auto& x : stateTrajectory) tt.stateTrajectory.push_back(x.std());
    return vector_t(tt.getDesiredState(t));
  }
};

// ---- ocs2_core/control/ControllerBase.h, FeedforwardController.h, LinearController.h (data carriers only) ----------------------------------
class ControllerBase {
 public:
  virtual ~ControllerBase() = default;
  virtual ControllerBase* clone() const = 0;
  virtual void clear() = 0;
  virtual vector_t computeInput(scalar_t t, const vector_t& x) = 0;
};
class FeedforwardController : public ControllerBase {
 public:
  FeedforwardController(scalar_array_t t, vector_array_t uff) : timeStamp_(std::move(t)), uffArray_(std::move(uff)) {}
  FeedforwardController* clone() const override { return new FeedforwardController(*this); }
  void clear() override {
    timeStamp_.clear();
    uffArray_.clear();
  }
  vector_t computeInput(scalar_t t, const vector_t&) override {
    std::vector<std::vector<double>> u;
    for (const auto& v : uffArray_) u.push_back(v.std());
    return vector_t(b200sqp::host::linearInterpolate(t, timeStamp_, u));
  }
  scalar_array_t timeStamp_;
  vector_array_t uffArray_;
};
class LinearController : public ControllerBase {  // u = uff + K x
 public:
  LinearController(scalar_array_t t, vector_array_t bias, matrix_array_t gain) : timeStamp_(std::move(t)), biasArray_(std::move(bias)), gainArray_(std::move(gain)) {}
  LinearController* clone() const override { return new LinearController(*this); }
  void clear() override {
    timeStamp_.clear();
    biasArray_.clear();
    gainArray_.clear();
  }
  vector_t computeInput(scalar_t t, const vector_t& x) override {
    std::size_t i = 0;
    while (i + 1 < timeStamp_.size() && timeStamp_[i + 1] <= t) ++i;
    vector_t u = biasArray_[i];
    for (std::size_t r = 0; r < u.size(); ++r)
      for (std::size_t c = 0; c < x.size(); ++c) u[r] += gainArray_[i](r, c) * x[c];
    return u;
  }
  scalar_array_t timeStamp_;
  vector_array_t biasArray_;
  matrix_array_t gainArray_;
};

// ---- ocs2_core/initialization/Initializer.h:44-66 ----------------------------------------------------------------------------------------------
class Initializer {
 public:
  virtual ~Initializer() = default;
  virtual Initializer* clone() const = 0;
  virtual void compute(scalar_t time, const vector_t& state, scalar_t nextTime, vector_t& input, vector_t& nextState) = 0;
};

// ---- ocs2_oc/oc_data/PerformanceIndex.h:42-98, PrimalSolution.h:43-106, TimeDiscretization.h:40-81 ---------------------------------------------
struct PerformanceIndex {
  scalar_t merit = 0.0, cost = 0.0, dualFeasibilitiesSSE = 0.0, dynamicsViolationSSE = 0.0, equalityConstraintsSSE = 0.0, inequalityConstraintsSSE = 0.0,
           equalityLagrangian = 0.0, inequalityLagrangian = 0.0;
};
struct PrimalSolution {
  PrimalSolution() = default;
  PrimalSolution(const PrimalSolution& o)
      : timeTrajectory_(o.timeTrajectory_), stateTrajectory_(o.stateTrajectory_), inputTrajectory_(o.inputTrajectory_), postEventIndices_(o.postEventIndices_),
        modeSchedule_(o.modeSchedule_), controllerPtr_(o.controllerPtr_ ? o.controllerPtr_->clone() : nullptr) {}
  PrimalSolution& operator=(const PrimalSolution& o) {
    PrimalSolution t(o);
    swap(t);
    return *this;
  }
  PrimalSolution(PrimalSolution&&) noexcept = default;
  PrimalSolution& operator=(PrimalSolution&&) noexcept = default;
  void swap(PrimalSolution& o) {
    timeTrajectory_.swap(o.timeTrajectory_);
    stateTrajectory_.swap(o.stateTrajectory_);
    inputTrajectory_.swap(o.inputTrajectory_);
    postEventIndices_.swap(o.postEventIndices_);
    std::swap(modeSchedule_, o.modeSchedule_);
    controllerPtr_.swap(o.controllerPtr_);
  }
  void clear() { *this = PrimalSolution(); }
  scalar_array_t timeTrajectory_;
  vector_array_t stateTrajectory_, inputTrajectory_;
  size_array_t postEventIndices_;
  ModeSchedule modeSchedule_;
  std::unique_ptr<ControllerBase> controllerPtr_;
};
struct AnnotatedTime {
  enum class Event { None, PreEvent, PostEvent };
  scalar_t time;
  Event event;
};
inline std::vector<AnnotatedTime> timeDiscretizationWithEvents(scalar_t initTime, scalar_t finalTime, scalar_t dt, const scalar_array_t& eventTimes) {
  std::vector<AnnotatedTime> out;
  for (const auto& a : b200sqp::host::timeDiscretizationWithEvents(initTime, finalTime, dt, eventTimes))
    out.push_back({a.time, a.event == b200sqp::host::EV_PRE ? AnnotatedTime::Event::PreEvent
                                                                : (a.event == b200sqp::host::EV_POST ? AnnotatedTime::Event::PostEvent : AnnotatedTime::Event::None)});
  return out;
}
inline scalar_t getIntervalStart(const AnnotatedTime& t) { return t.time + (t.event == AnnotatedTime::Event::PostEvent ? 1e-9 : 0.0); }
inline scalar_t getIntervalEnd(const AnnotatedTime& t) { return t.time - (t.event == AnnotatedTime::Event::PreEvent ? 1e-9 : 0.0); }

// ---- ocs2_oc/synchronized_module/ReferenceManagerInterface.h:44-75 -------------------------------------------------------------------------------
class ReferenceManagerInterface {
 public:
  virtual ~ReferenceManagerInterface() = default;
  virtual void preSolverRun(scalar_t initTime, scalar_t finalTime, const vector_t& initState, std::size_t initMode) = 0;
  virtual const ModeSchedule& getModeSchedule() const = 0;
  virtual const TargetTrajectories& getTargetTrajectories() const = 0;
};

// ---- ocs2_oc/oc_solver/SolverBase.h:52-273 -------------------------------------------------------------------------------------------------------
class SolverBase {
 public:
  virtual ~SolverBase() = default;
  virtual void reset() = 0;
  void run(scalar_t initTime, const vector_t& initState, std::size_t initMode, scalar_t finalTime) {
    referenceManagerPtr_->preSolverRun(initTime, finalTime, initState, initMode);   // preRun (SolverBase.cpp:88-100)
    runImpl(initTime, initState, finalTime);
  }
  void setReferenceManager(std::shared_ptr<ReferenceManagerInterface> p) {
    if (!p) throw std::runtime_error("[SolverBase] ReferenceManager pointer cannot be a nullptr!");
    referenceManagerPtr_ = std::move(p);
  }
  ReferenceManagerInterface& getReferenceManager() { return *referenceManagerPtr_; }
  const ReferenceManagerInterface& getReferenceManager() const { return *referenceManagerPtr_; }
  virtual const OptimalControlProblem& getOptimalControlProblem() const = 0;
  virtual const PerformanceIndex& getPerformanceIndeces() const = 0;
  virtual std::size_t getNumIterations() const = 0;
  virtual const std::vector<PerformanceIndex>& getIterationsLog() const = 0;
  virtual scalar_t getFinalTime() const = 0;
  virtual void getPrimalSolution(scalar_t finalTime, PrimalSolution* primalSolutionPtr) const = 0;
  PrimalSolution primalSolution(scalar_t finalTime) const {
    PrimalSolution p;
    getPrimalSolution(finalTime, &p);
    return p;
  }
  virtual const DualSolution* getDualSolution() const { return nullptr; }
  virtual const ProblemMetrics& getSolutionMetrics() const = 0;
  virtual ScalarFunctionQuadraticApproximation getValueFunction(scalar_t time, const vector_t& state) const = 0;
  virtual ScalarFunctionQuadraticApproximation getHamiltonian(scalar_t time, const vector_t& state, const vector_t& input) = 0;
  virtual vector_t getStateInputEqualityConstraintLagrangian(scalar_t time, const vector_t& state) const = 0;
  virtual MultiplierCollection getIntermediateDualSolution(scalar_t time) const = 0;
  virtual std::string getBenchmarkingInfo() const { return {}; }

 private:
  virtual void runImpl(scalar_t initTime, const vector_t& initState, scalar_t finalTime) = 0;
  virtual void runImpl(scalar_t initTime, const vector_t& initState, scalar_t finalTime, const ControllerBase* externalControllerPtr) = 0;
  virtual void runImpl(scalar_t initTime, const vector_t& initState, scalar_t finalTime, const PrimalSolution& primalSolution) = 0;
  std::shared_ptr<ReferenceManagerInterface> referenceManagerPtr_;
};

// ---- ocs2_oc/multiple_shooting/Initialization.h:50-120 (Initialization.cpp:35-79) ----------------------------------------------------------------
namespace multiple_shooting {
inline void initializeStateInputTrajectories(const vector_t& initState, const std::vector<AnnotatedTime>& time, const PrimalSolution& primal,
                                             Initializer& initializer, vector_array_t& x, vector_array_t& u) {
  const int N = static_cast<int>(time.size()) - 1;
  x.clear();
  u.clear();
  const bool warm = primal.timeTrajectory_.size() >= 2;
  const scalar_t tStateTill = warm ? primal.timeTrajectory_.back() : time[0].time;
  const scalar_t tInputTill = warm ? primal.timeTrajectory_[primal.timeTrajectory_.size() - 2] : time[0].time;
  auto interp = [](scalar_t t, const scalar_array_t& ts, const vector_array_t& vs) {
    std::vector<std::vector<double>> v;
    for (const auto& a : vs) v.push_back(a.std());
    return vector_t(b200sqp::host::linearInterpolate(t, ts, v));
  };
  const scalar_t tInit = getIntervalStart(time[0]);
  x.push_back(tInit < tStateTill ? interp(tInit, primal.timeTrajectory_, primal.stateTrajectory_) : initState);
  for (int i = 0; i < N; ++i) {
    if (time[i].event == AnnotatedTime::Event::PreEvent) {
      u.push_back(vector_t());   // no input at a pre-event node
      x.push_back(x.back());
      continue;
    }
    const scalar_t t = getIntervalStart(time[i]), tNext = getIntervalEnd(time[i + 1]);
    if (t > tInputTill || tNext > tStateTill) {
      vector_t ui, xn;
      initializer.compute(t, x.back(), tNext, ui, xn);
      u.push_back(ui);
      x.push_back(xn);
    } else {
      u.push_back(interp(t, primal.timeTrajectory_, primal.inputTrajectory_));
      x.push_back(interp(tNext, primal.timeTrajectory_, primal.stateTrajectory_));
    }
  }
}
}  // namespace multiple_shooting

// ---- ocs2_sqp/SqpSettings.h:40-87 ----------------------------------------------------------------------------------------------------------------
namespace sqp {
struct Settings {
  std::size_t sqpIteration = 10;
  scalar_t deltaTol = 1e-6, costTol = 1e-4, alpha_decay = 0.5, alpha_min = 1e-4, g_max = 1e6, g_min = 1e-6, armijoFactor = 1e-4, gamma_c = 1e-6;
  bool useFeedbackPolicy = true, createValueFunction = false;
  struct {
    scalar_t reg_prim = 1e-12;
  } hpipmSettings;   // hpipm_interface::Settings::reg_prim (HpipmInterfaceSettings.h:45-56)
  scalar_t dt = 0.01;
  bool projectStateInputEqualityConstraints = true, extractProjectionMultiplier = false;
  bool printSolverStatus = false, printSolverStatistics = false, printLinesearch = false, enableLogging = true;
  std::size_t nThreads = 4;
};
}  // namespace sqp

// ---- humanoid_common_mpc: common/Types.h (contact_flag_t), swing_foot_planner/SwingTrajectoryPlanner.h:62-68,
//      reference_manager/SwitchedModelReferenceManager.h:47-100 ---------------------------------------------------------------------------------------
namespace humanoid {
using contact_flag_t = std::array<bool, 2>;
class SwingTrajectoryPlanner {
 public:
  explicit SwingTrajectoryPlanner(const b200sqp::host::SwingTrajectoryPlanner* p) : p_(p) {}
  scalar_t getZpositionConstraint(std::size_t leg, scalar_t t) const { return p_->zReference(static_cast<int>(leg), t)[0]; }
  scalar_t getZvelocityConstraint(std::size_t leg, scalar_t t) const { return p_->zReference(static_cast<int>(leg), t)[1]; }
  scalar_t getZaccelerationConstraint(std::size_t leg, scalar_t t) const { return p_->zReference(static_cast<int>(leg), t)[2]; }
  scalar_t getImpactProximityFactor(std::size_t leg, scalar_t t) const { return p_->impactProximityFactor(static_cast<int>(leg), t); }

 private:
  const b200sqp::host::SwingTrajectoryPlanner* p_;
};
class SwitchedModelReferenceManager : public ReferenceManagerInterface {
 public:
  explicit SwitchedModelReferenceManager(const b200sqp::host::HostModel& m) : host_(m), planner_(std::make_shared<SwingTrajectoryPlanner>(&host_.getSwingTrajectoryPlanner())) {}
  b200sqp::host::SwitchedModelReferenceManager& host() { return host_; }
  void setTargetTrajectories(TargetTrajectories tt) {
    b200sqp::host::TargetTrajectories h;
    h.timeTrajectory = tt.timeTrajectory;
    for (const auto& x : tt.stateTrajectory) h.stateTrajectory.push_back(x.std());
    host_.setTargetTrajectories(std::move(h));
    targets_ = std::move(tt);
  }
  void preSolverRun(scalar_t initTime, scalar_t finalTime, const vector_t&, std::size_t) override {
    host_.preSolverRun(initTime, finalTime);
    modeSchedule_.eventTimes = host_.getModeSchedule().eventTimes;
    modeSchedule_.modeSequence.assign(host_.getModeSchedule().modeSequence.begin(), host_.getModeSchedule().modeSequence.end());
  }
  const ModeSchedule& getModeSchedule() const override { return modeSchedule_; }
  const TargetTrajectories& getTargetTrajectories() const override { return targets_; }
  contact_flag_t getContactFlags(scalar_t t) const {
    const auto c = b200sqp::host::modeNumber2StanceLeg(host_.getModeSchedule().modeAtTime(t));
    return {static_cast<bool>(c[0]), static_cast<bool>(c[1])};
  }
  const std::shared_ptr<SwingTrajectoryPlanner>& getSwingTrajectoryPlanner() const { return planner_; }
  scalar_t getPhaseVariable(scalar_t t) const { return b200sqp::host::getPhaseVariable(host_.getModeSchedule(), t); }

 private:
  b200sqp::host::SwitchedModelReferenceManager host_;
  std::shared_ptr<SwingTrajectoryPlanner> planner_;
  ModeSchedule modeSchedule_;
  TargetTrajectories targets_;
};
}  // namespace humanoid
}  // namespace ocs2
