// Compiles include/ocs2_sqp/B200SqpSolver.h against the stand-in ocs2 headers and drives it like SqpMpc::calculateController does
// (SqpMpc.h:49-66): solver.run(t0, x0, initMode, tf) twice (cold start, then warm start from its own solution), and compares the first
// solution with b200sqp::host::SqpSolver on the same instance.  Exit code 0 = agreement; 3 = no CUDA device (the shim threw, as it must).
#include <cstdio>
#include <cstring>

#include <ocs2_sqp/B200SqpSolver.h>

#include "../../wb_humanoid_mpc_b200/host/SqpSolver.hpp"

using namespace ocs2;

class WeightCompInit final : public Initializer {   // humanoid_common_mpc/initialization/WeightCompInitializer.cpp:66-70
 public:
  explicit WeightCompInit(const b200sqp::host::HostModel& m, const humanoid::SwitchedModelReferenceManager* rm) : m_(&m), rm_(rm) {}
  WeightCompInit* clone() const override { return new WeightCompInit(*this); }
  void compute(scalar_t time, const vector_t& state, scalar_t, vector_t& input, vector_t& nextState) override {
    const auto c = rm_->getContactFlags(time);
    input = vector_t(b200sqp::host::weightCompensatingInput(*m_, c[0], c[1]));
    nextState = state;
  }

 private:
  const b200sqp::host::HostModel* m_;
  const humanoid::SwitchedModelReferenceManager* rm_;
};

int main(int argc, char** argv) {
  const std::string file = argc > 1 ? argv[1] : "wb_humanoid_mpc_b200/data/g1_wb_model.txt";
  namespace h = b200sqp::host;
  h::HostModel model = h::loadModelFile(file);
  sqp::Settings s;
  s.sqpIteration = 1;
  s.dt = model.dt;
  s.deltaTol = model.sqpSettings.delta_tol;
  s.costTol = model.sqpSettings.cost_tol;
  s.g_max = model.sqpSettings.g_max;
  s.g_min = model.sqpSettings.g_min;
  s.useFeedbackPolicy = false;
  auto rm = std::make_shared<humanoid::SwitchedModelReferenceManager>(model);
  const double t0 = 0.0, tf = 1.1;
  vector_t x0(model.initialState);
  x0[2] = model.defaultBaseHeight;
  rm->host().setGait("walk", 0.0, 10.0);
  const auto tt = h::commandedVelocityToTargetTrajectories(model, t0, x0.std(), {0.4, 0.0, model.defaultBaseHeight, 0.1}, tf - t0);
  TargetTrajectories targets;
  targets.timeTrajectory = tt.timeTrajectory;
  for (const auto& x : tt.stateTrajectory) targets.stateTrajectory.push_back(vector_t(x));
  rm->setTargetTrajectories(targets);
  WeightCompInit init(model, rm.get());
  OptimalControlProblem ocp;
  std::unique_ptr<B200SqpSolver> solver;
  try {
    solver.reset(new B200SqpSolver(s, ocp, init, model.desc));
  } catch (const std::runtime_error& e) {
    std::printf("SHIM_NO_DEVICE %s\n", e.what());
    return std::strstr(e.what(), "no CUDA device") ? 3 : 4;
  }
  solver->setReferenceManager(rm);
  bool threw = false;
  try {
    solver->getIterationsLog();
  } catch (const std::runtime_error&) {
    threw = true;   // empty log throws, as in the reference
  }
  if (!threw) return 5;
  solver->run(t0, x0, 3, tf);
  const PrimalSolution p1 = solver->primalSolution(tf);
  // the same instance through the batched host layer
  h::SqpSolver ref(model, model.sqpSettings, 1, 0);
  ref.getReferenceManager(0).setGait("walk", 0.0, 10.0);
  ref.getReferenceManager(0).setTargetTrajectories(tt);
  ref.run(t0, {x0.std()}, tf);
  const h::PrimalSolution& q = ref.primalSolution(0);
  if (q.timeTrajectory_.size() != p1.timeTrajectory_.size()) return 6;
  double err = 0.0;
  for (size_t i = 0; i < p1.timeTrajectory_.size(); ++i) {
    for (size_t k = 0; k < p1.stateTrajectory_[i].size(); ++k) err = std::max(err, std::fabs(p1.stateTrajectory_[i][k] - q.stateTrajectory_[i][k]));
    for (size_t k = 0; k < p1.inputTrajectory_[i].size(); ++k) err = std::max(err, std::fabs(p1.inputTrajectory_[i][k] - q.inputTrajectory_[i][k]));
  }
  // second MPC cycle: warm start from the shim's own primal solution
  const double t1 = 1.0 / 60.0;
  vector_t x1 = p1.controllerPtr_ ? x0 : x0;
  for (size_t k = 0; k < x1.size(); ++k) x1[k] = p1.stateTrajectory_[0][k] + (p1.stateTrajectory_[1][k] - p1.stateTrajectory_[0][k]) * (t1 / (p1.timeTrajectory_[1] - p1.timeTrajectory_[0]));
  solver->run(t1, x1, 3, t1 + (tf - t0));
  const auto& log = solver->getIterationsLog();
  const auto ms = solver->getBenchmarks();
  std::printf("SHIM_OK nodes=%zu max_abs_diff_vs_host_layer=%.3e second_run_merit=%.6f iterations=%zu lq_ms=%.3f\n", p1.timeTrajectory_.size(), err, log.back().merit,
              solver->getNumIterations(), ms[0]);
  return err < 1e-9 ? 0 : 7;
}
