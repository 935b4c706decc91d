// Host layer (C++17, no dependencies): the wire format of an MPC policy, ocs2_msgs::mpc_flattened_controller, as plain structs
// (SURVEY.md section 8(f)-4).  Restates MPC_ROS_Interface::createMpcPolicyMsg (lib/ocs2_ros2/ocs2_ros_interfaces/src/mpc/MPC_ROS_Interface.cpp:98-178)
// and the controllers' flatten / unFlatten (ocs2_core/src/control/FeedforwardController.cpp:96-150, LinearController.cpp:92-190): a node that
// publishes /humanoid/mpc_policy copies these fields into the ROS message one to one; MRT_ROS_Interface::readPolicyMsg reads them back.
//   FEEDFORWARD: data[k] = uff(t_k)                         (nu floats)
//   LINEAR:      data[k] = [uff[0], K[0, :], uff[1], K[1, :], ...]   (nu * (1 + nx) floats, row-major per input)
#pragma once
#include <cstdint>
#include <stdexcept>
#include <vector>

#include "references.hpp"

namespace b200sqp::host {

struct MpcFlattenedController {
  static constexpr uint8_t CONTROLLER_FEEDFORWARD = 0, CONTROLLER_LINEAR = 1;   // mpc_flattened_controller.msg
  uint8_t controllerType = CONTROLLER_FEEDFORWARD;
  double initTime = 0.0;                      // initObservation.time
  std::vector<float> initState, initInput;    // initObservation.state / .input (mpc_observation uses float32 arrays)
  std::vector<double> eventTimes;             // modeSchedule
  std::vector<int8_t> modeSequence;
  std::vector<double> timeTrajectory;
  std::vector<uint16_t> postEventIndices;
  std::vector<std::vector<float>> stateTrajectory, inputTrajectory, data;
};

// K: remapped gains of the primal solution's nodes, [n - 1][nu * nx] column-major (b200sqp_download), or nullptr for the feed-forward policy
inline MpcFlattenedController createMpcPolicyMsg(const PrimalSolution& p, double initTime, const vector_t& initState, const double* K, int nx, int nu) {
  MpcFlattenedController m;
  const size_t n = p.timeTrajectory_.size();
  if (p.stateTrajectory_.size() != n || p.inputTrajectory_.size() != n) throw std::runtime_error("createMpcPolicyMsg: inconsistent primal solution");
  m.controllerType = K ? MpcFlattenedController::CONTROLLER_LINEAR : MpcFlattenedController::CONTROLLER_FEEDFORWARD;
  m.initTime = initTime;
  m.initState.assign(initState.begin(), initState.end());
  m.initInput.assign(static_cast<size_t>(nu), 0.0f);
  m.eventTimes = p.modeSchedule_.eventTimes;
  m.modeSequence.assign(p.modeSchedule_.modeSequence.begin(), p.modeSchedule_.modeSequence.end());
  m.timeTrajectory = p.timeTrajectory_;
  for (size_t k = 0; k < n; ++k)
    if (p.postEventIndices_[k] == EV_POST) m.postEventIndices.push_back(static_cast<uint16_t>(k));   // node annotations -> indices
  for (size_t k = 0; k < n; ++k) {
    m.stateTrajectory.emplace_back(p.stateTrajectory_[k].begin(), p.stateTrajectory_[k].end());
    m.inputTrajectory.emplace_back(p.inputTrajectory_[k].begin(), p.inputTrajectory_[k].end());
    std::vector<float> d;
    if (!K) {
      d.assign(p.inputTrajectory_[k].begin(), p.inputTrajectory_[k].end());
    } else {
      // uff = u - K x (SqpSolver.cpp:338-340); the terminal sample repeats the last stage's gain like the last input
      const size_t src = (k + 1 < n) ? k : n - 2;
      const double* Kk = K + src * static_cast<size_t>(nu) * nx;
      d.reserve(static_cast<size_t>(nu) * (1 + nx));
      for (int i = 0; i < nu; ++i) {
        double uff = p.inputTrajectory_[k][i];
        for (int j = 0; j < nx; ++j) uff -= Kk[i + nu * j] * p.stateTrajectory_[k][j];
        d.push_back(static_cast<float>(uff));
        for (int j = 0; j < nx; ++j) d.push_back(static_cast<float>(Kk[i + nu * j]));
      }
    }
    m.data.push_back(std::move(d));
  }
  return m;
}

// LinearController::unFlatten / FeedforwardController::unFlatten + computeInput at sample k (what MRT_ROS_Interface evaluates)
inline std::vector<double> evaluatePolicySample(const MpcFlattenedController& m, size_t k, const vector_t& x) {
  const std::vector<float>& d = m.data.at(k);
  if (m.controllerType == MpcFlattenedController::CONTROLLER_FEEDFORWARD) return std::vector<double>(d.begin(), d.end());
  const size_t nx = x.size(), nu = d.size() / (1 + nx);
  std::vector<double> u(nu);
  for (size_t i = 0; i < nu; ++i) {
    double s = d[i * (1 + nx)];
    for (size_t j = 0; j < nx; ++j) s += static_cast<double>(d[i * (1 + nx) + 1 + j]) * x[j];
    u[i] = s;
  }
  return u;
}

}  // namespace b200sqp::host
