// Host layer: what SolverBase::preRun + SqpSolver::runImpl's bookkeeping hand to the LQ stage, as flat per-node arrays in the layout of
// b200sqp_upload_instances.  C++ restatement (names follow the reference) of
//   GaitSchedule                 humanoid_nmpc/humanoid_common_mpc/src/gait/GaitSchedule.cpp:46-139
//   ModeSchedule::modeAtTime     lib/ocs2_ros2/ocs2_core/src/reference/ModeSchedule.cpp:48-51
//   SwitchedModelReferenceManager::modifyReferences / getPhaseVariable / getContactFlags
//                                humanoid_nmpc/humanoid_common_mpc/src/reference_manager/SwitchedModelReferenceManager.cpp:54-154
//   SwingTrajectoryPlanner / SplineCpg / CubicSpline
//                                humanoid_nmpc/humanoid_common_mpc/src/swing_foot_planner/SwingTrajectoryPlanner.cpp:50-271,
//                                SplineCpg.cpp:38-63, CubicSpline.cpp:38-85
//   timeDiscretizationWithEvents lib/ocs2_ros2/ocs2_oc/src/oc_data/TimeDiscretization.cpp:40-114
//   commandedVelocityToTargetTrajectories
//                                humanoid_nmpc/humanoid_wb_mpc/src/command/WBMpcTargetTrajectoriesCalculator.cpp:82-136
//   WeightCompInitializer, initializeStateInputTrajectories, toPrimalSolution, LinearInterpolation
//                                humanoid_common_mpc/src/initialization/WeightCompInitializer.cpp:66-70,
//                                ocs2_oc/src/multiple_shooting/Initialization.cpp:35-79, Helpers.cpp:60-82,
//                                ocs2_core/include/ocs2_core/misc/implementation/LinearInterpolation.h:67-129
// tests/test_host_cpp.py checks every array against the independent Python restatement (wb_humanoid_mpc_b200/references.py).
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>

#include "model_file.hpp"

namespace b200sqp::host {

using vector_t = std::vector<double>;
constexpr double kLimitEps = 1e-6;   // numeric_traits::limitEpsilon
constexpr double kWeakEps = 1e-9;    // numeric_traits::weakEpsilon
enum Event : uint8_t { EV_NONE = 0, EV_PRE = 1, EV_POST = 2 };  // AnnotatedTime::Event

inline std::array<bool, 2> modeNumber2StanceLeg(int mode) {  // {left, right}
  static const bool T[4][2] = {{false, false}, {false, true}, {true, false}, {true, true}};
  return {T[mode][0], T[mode][1]};
}
inline size_t lowerBoundIndex(const vector_t& v, double t) { return static_cast<size_t>(std::lower_bound(v.begin(), v.end(), t) - v.begin()); }
inline size_t upperBoundIndex(const vector_t& v, double t) { return static_cast<size_t>(std::upper_bound(v.begin(), v.end(), t) - v.begin()); }

struct ModeSchedule {
  vector_t eventTimes;
  std::vector<int> modeSequence{STANCE};
  int modeAtTime(double t) const { return modeSequence[lowerBoundIndex(eventTimes, t)]; }
};

class GaitSchedule {
 public:
  GaitSchedule() : ms_{{0.5}, {STANCE, STANCE}}, tModes_{STANCE}, tTimes_{0.0, 0.5} {}
  void insertModeSequenceTemplate(const GaitTemplate& g, double startTime, double finalTime) {
    tModes_ = g.modeSequence;
    tTimes_ = g.switchingTimes;
    auto& ev = ms_.eventTimes;
    auto& seq = ms_.modeSequence;
    const size_t idx = lowerBoundIndex(ev, startTime);
    if (idx < ev.size()) {
      ev.erase(ev.begin() + idx, ev.end());
      seq.erase(seq.begin() + idx + 1, seq.end());
    }
    const double ptst = (!seq.empty() && seq.back() == STANCE) ? 0.0 : phaseTransitionStanceTime_;
    if (ptst > 0.0) {
      ev.push_back(startTime);
      seq.push_back(STANCE);
    }
    tileModeSequenceTemplate(startTime + ptst, finalTime);
  }
  ModeSchedule getModeSchedule(double lowerBoundTime, double upperBoundTime) {
    auto& ev = ms_.eventTimes;
    auto& seq = ms_.modeSequence;
    const size_t idx = lowerBoundIndex(ev, lowerBoundTime);
    if (idx > 0) {
      ev.erase(ev.begin(), ev.begin() + (idx - 1));
      seq.erase(seq.begin(), seq.begin() + (idx - 1));
      seq[0] = STANCE;
    }
    const double tilingStart = ev.empty() ? upperBoundTime : ev.back();
    if (!ev.empty()) ev.pop_back();
    if (!seq.empty()) seq.pop_back();
    tileModeSequenceTemplate(tilingStart, upperBoundTime);
    return ms_;
  }

 private:
  void tileModeSequenceTemplate(double startTime, double finalTime) {
    auto& ev = ms_.eventTimes;
    auto& seq = ms_.modeSequence;
    if (tModes_.empty()) return;
    if (!ev.empty() && startTime <= ev.back()) throw std::runtime_error("The initial time for template-tiling is not greater than the last event time.");
    ev.push_back(startTime);
    while (ev.back() < finalTime)
      for (size_t i = 0; i < tModes_.size(); ++i) {
        seq.push_back(tModes_[i]);
        ev.push_back(ev.back() + (tTimes_[i + 1] - tTimes_[i]));
      }
    seq.push_back(STANCE);
  }
  ModeSchedule ms_;
  std::vector<int> tModes_;
  vector_t tTimes_;
  double phaseTransitionStanceTime_ = 0.0;
};

struct SplineNode {
  double time, position, velocity;
};
class CubicSpline {
 public:
  CubicSpline() = default;
  CubicSpline(SplineNode s, SplineNode e) : t0_(s.time), dt_(e.time - s.time) {
    const double dp = e.position - s.position, dv = e.velocity - s.velocity;
    c0_ = s.position;
    c1_ = s.velocity * dt_;
    c2_ = -(3.0 * s.velocity + dv) * dt_ + 3.0 * dp;
    c3_ = (2.0 * s.velocity + dv) * dt_ - 2.0 * dp;
  }
  double position(double t) const {
    const double tn = (t - t0_) / dt_;
    return c3_ * tn * tn * tn + c2_ * tn * tn + c1_ * tn + c0_;
  }
  double velocity(double t) const {
    const double tn = (t - t0_) / dt_;
    return (3.0 * c3_ * tn * tn + 2.0 * c2_ * tn + c1_) / dt_;
  }
  double acceleration(double t) const {
    const double tn = (t - t0_) / dt_;
    return (6.0 * c3_ * tn + 2.0 * c2_) / (dt_ * dt_);
  }

 private:
  double t0_ = 0, dt_ = 1, c0_ = 0, c1_ = 0, c2_ = 0, c3_ = 0;
};
class SplineCpg {
 public:
  SplineCpg(SplineNode liftOff, double midHeight, SplineNode touchDown)
      : mid_((liftOff.time + touchDown.time) / 2), left_(liftOff, SplineNode{mid_, midHeight, 0.0}), right_(SplineNode{mid_, midHeight, 0.0}, touchDown) {}
  double position(double t) const { return t < mid_ ? left_.position(t) : right_.position(t); }
  double velocity(double t) const { return t < mid_ ? left_.velocity(t) : right_.velocity(t); }
  double acceleration(double t) const { return t < mid_ ? left_.acceleration(t) : right_.acceleration(t); }

 private:
  double mid_;
  CubicSpline left_, right_;
};

class SwingTrajectoryPlanner {
 public:
  explicit SwingTrajectoryPlanner(SwingTrajectoryConfig cfg, int numFeet = 2) : c_(cfg), n_(numFeet) {}
  void update(const ModeSchedule& ms, double terrainHeight = 0.0) {
    const auto& modes = ms.modeSequence;
    const auto& ev = ms.eventTimes;
    const int nph = static_cast<int>(modes.size());
    const double liftH = terrainHeight, touchH = terrainHeight + c_.touchDownHeightOffset;
    events_ = ev;
    height_.assign(n_, {});
    impact_.assign(n_, {});
    for (int j = 0; j < n_; ++j) {
      std::vector<bool> flags(nph);
      for (int p = 0; p < nph; ++p) flags[p] = modeNumber2StanceLeg(modes[p])[j];
      for (int p = 0; p < nph; ++p) {
        if (flags[p]) {
          height_[j].emplace_back(SplineNode{0.0, liftH, 0.0}, liftH, SplineNode{1.0, liftH, 0.0});
          impact_[j].emplace_back(SplineNode{0.0, 1.0, 0.0}, 1.0, SplineNode{1.0, 1.0, 0.0});
          continue;
        }
        int start = -1, fin = nph - 1;
        for (int ip = p - 1; ip >= 0; --ip)
          if (flags[ip]) {
            start = ip;
            break;
          }
        for (int ip = p + 1; ip < nph; ++ip)
          if (flags[ip]) {
            fin = ip - 1;
            break;
          }
        if (start < 0) throw std::runtime_error("The time of take-off for the first swing of the EE with ID " + std::to_string(j) + " is not defined.");
        if (fin >= nph - 1) throw std::runtime_error("The time of touch-down for the last swing of the EE with ID " + std::to_string(j) + " is not defined.");
        const double ts = ev[start], tf = ev[fin];
        const bool prevC = flags[p - 1], nextC = flags[p + 1];
        const double midV = c_.ipfMidPointValue;
        if (prevC && nextC) {
          const double s = std::min(1.0, (tf - ts) / c_.swingTimeScale);
          height_[j].emplace_back(SplineNode{ts, liftH, s * c_.liftOffVelocity}, std::min(liftH, touchH) + s * c_.swingHeight,
                                  SplineNode{tf, touchH, s * c_.touchDownVelocity});
          impact_[j].emplace_back(SplineNode{ts, 1.0, s * c_.ipfLiftOffVelocity}, midV, SplineNode{tf, 1.0, s * c_.ipfTouchDownVelocity});
        } else if (prevC) {
          const double mid = liftH + c_.swingHeight;
          height_[j].emplace_back(SplineNode{ts, liftH, c_.liftOffVelocity}, mid, SplineNode{tf, mid, 0.0});
          impact_[j].emplace_back(SplineNode{ts, 1.0, c_.ipfLiftOffVelocity}, midV, SplineNode{tf, midV, 0.0});
        } else if (nextC) {
          const double mid = touchH + c_.swingHeight;
          height_[j].emplace_back(SplineNode{ts, mid, 0.0}, mid, SplineNode{tf, touchH, c_.touchDownVelocity});
          impact_[j].emplace_back(SplineNode{ts, midV, 0.0}, midV, SplineNode{tf, 1.0, c_.ipfTouchDownVelocity});
        } else {
          const double mid = touchH + c_.swingHeight;
          height_[j].emplace_back(SplineNode{ts, mid, 0.0}, mid, SplineNode{tf, mid, 0.0});
          impact_[j].emplace_back(SplineNode{ts, midV, 0.0}, midV, SplineNode{tf, midV, 0.0});
        }
      }
    }
  }
  // getZpositionConstraint / getZvelocityConstraint / getZaccelerationConstraint
  std::array<double, 3> zReference(int leg, double t) const {
    const SplineCpg& s = height_[leg][lowerBoundIndex(events_, t)];
    return {s.position(t), s.velocity(t), s.acceleration(t)};
  }
  double impactProximityFactor(int leg, double t) const { return impact_[leg][lowerBoundIndex(events_, t)].position(t); }

 private:
  SwingTrajectoryConfig c_;
  int n_;
  vector_t events_;
  std::vector<std::vector<SplineCpg>> height_, impact_;
};

// SwitchedModelReferenceManager::getPhaseVariable
inline double getPhaseVariable(const ModeSchedule& ms, double t) {
  const auto& ev = ms.eventTimes;
  const size_t it = upperBoundIndex(ev, t);
  // the reference dereferences both neighbours unchecked (SwitchedModelReferenceManager.cpp:63-65); outside the event range the phase is a
  // STANCE phase and only `prv` is used, for a mode lookup: any time before the first event gives the same answer
  const double nxt = it < ev.size() ? ev[it] : ev.back();
  const double prv = it > 0 ? ev[it - 1] : ev.front() - 1.0;
  const int m = ms.modeAtTime(t);
  if (m == LF) return 0.5 * (t - prv) / (nxt - prv);
  if (m == RF) return 0.5 + 0.5 * (t - prv) / (nxt - prv);
  return ms.modeAtTime(prv - 0.01) == LF ? 0.5 : 0.0;
}

struct AnnotatedTime {
  double time;
  Event event;
};
inline std::vector<AnnotatedTime> timeDiscretizationWithEvents(double initTime, double finalTime, double dt, const vector_t& eventTimes,
                                                               double dtMin = 10.0 * kLimitEps) {
  std::vector<AnnotatedTime> td{{initTime, EV_NONE}};
  size_t nextIdx = lowerBoundIndex(eventTimes, initTime);
  AnnotatedTime next{initTime, EV_NONE};
  while (td.back().time < finalTime) {
    next = {next.time + dt, EV_NONE};
    if (nextIdx < eventTimes.size() && next.time >= eventTimes[nextIdx]) {
      next = {eventTimes[nextIdx], EV_PRE};
      ++nextIdx;
    }
    if (next.time >= finalTime) next = {finalTime, EV_NONE};
    if (next.time > td.back().time + dtMin) td.push_back(next);
    else td.back() = next;
  }
  if (td.front().event == EV_PRE) td.front().event = EV_POST;
  std::vector<AnnotatedTime> out;
  for (const auto& a : td) {
    out.push_back(a);
    if (a.event == EV_PRE) out.push_back({a.time, EV_POST});
  }
  return out;
}
inline double getIntervalStart(const AnnotatedTime& a) { return a.time + (a.event == EV_POST ? kWeakEps : 0.0); }
inline double getIntervalEnd(const AnnotatedTime& a) { return a.time - (a.event == EV_PRE ? kWeakEps : 0.0); }

struct TargetTrajectories {
  vector_t timeTrajectory;
  std::vector<vector_t> stateTrajectory;
  // TargetTrajectories::getDesiredState: linear interpolation with clamping
  vector_t getDesiredState(double t) const {
    if (t <= timeTrajectory.front()) return stateTrajectory.front();
    if (t >= timeTrajectory.back()) return stateTrajectory.back();
    const size_t i = upperBoundIndex(timeTrajectory, t) - 1;
    const double a = (timeTrajectory[i + 1] - t) / (timeTrajectory[i + 1] - timeTrajectory[i]);
    vector_t x(stateTrajectory[i].size());
    for (size_t k = 0; k < x.size(); ++k) x[k] = a * stateTrajectory[i][k] + (1 - a) * stateTrajectory[i + 1][k];
    return x;
  }
};

// velocity command [v_x, v_y, pelvis height, yaw rate] -> 3-knot target trajectories (steady-state command filter)
inline TargetTrajectories commandedVelocityToTargetTrajectories(const HostModel& m, double initTime, const vector_t& x0, const std::array<double, 4>& cmd,
                                                                double horizon) {
  const int nj = m.nj, nv = 6 + nj;
  std::array<double, 6> pose{x0[0], x0[1], x0[2], x0[3], 0.0, 0.0};
  const double yaw = pose[3];
  const double vgx = std::cos(yaw) * cmd[0] - std::sin(yaw) * cmd[1], vgy = std::sin(yaw) * cmd[0] + std::cos(yaw) * cmd[1];
  const std::array<double, 6> baseVel{vgx, vgy, 0.0, cmd[3], 0.0, 0.0};
  const double tMid = 0.7 * horizon;
  const std::array<double, 3> avg{(x0[nv] + vgx) / 2, (x0[nv + 1] + vgy) / 2, (x0[nv + 5] + cmd[3]) / 2};
  pose[2] = cmd[2];
  auto integrate = [](std::array<double, 6> p, const std::array<double, 3>& av, double h, double dT) {
    p[0] += av[0] * dT;
    p[1] += av[1] * dT;
    p[2] = h;
    p[3] += av[2] * dT;
    p[4] = p[5] = 0.0;
    return p;
  };
  const auto mid = integrate(pose, avg, cmd[2], tMid);
  const auto fin = integrate(mid, {vgx, vgy, cmd[3]}, cmd[2], horizon - tMid);
  TargetTrajectories tt;
  tt.timeTrajectory = {initTime, initTime + tMid, initTime + horizon};
  for (const auto& p : {pose, mid, fin}) {
    vector_t x(m.nx, 0.0);
    for (int k = 0; k < 6; ++k) {
      x[k] = p[k];
      x[nv + k] = baseVel[k];
    }
    for (int j = 0; j < nj; ++j) x[6 + j] = m.defaultJointState[j];
    tt.stateTrajectory.push_back(x);
  }
  return tt;
}

// CentroidalMpcTargetTrajectoriesCalculator::commandedVelocityToTargetTrajectories (humanoid_centroidal_mpc/src/command/
// CentroidalMpcTargetTrajectoriesCalculator.cpp:82-160).  baseVel = Ab^-1 * x0[0..6) -- the inverse base block of the centroidal momentum
// matrix times the NORMALIZED momentum, as the reference computes it (centroidalBaseVelocity() below obtains it from the device); the yaw
// average uses baseVel[5], the roll rate in the ZYX ordering, as the reference does.
inline TargetTrajectories commandedVelocityToTargetTrajectoriesCentroidal(const HostModel& m, double initTime, const vector_t& x0,
                                                                          const std::array<double, 4>& cmd, double horizon,
                                                                          const std::array<double, 6>& baseVel) {
  const int nj = m.nj;
  std::array<double, 6> pose{x0[6], x0[7], x0[8], x0[9], 0.0, 0.0};
  const double yaw = pose[3];
  const double vgx = std::cos(yaw) * cmd[0] - std::sin(yaw) * cmd[1], vgy = std::sin(yaw) * cmd[0] + std::cos(yaw) * cmd[1];
  const std::array<double, 6> momentum{vgx, vgy, 0.0, 0.0, 0.0, cmd[3] / m.totalMass};
  const double tMid = 0.7 * horizon;
  const std::array<double, 3> avg{(baseVel[0] + vgx) / 2, (baseVel[1] + vgy) / 2, (baseVel[5] + cmd[3]) / 2};
  pose[2] = cmd[2];
  auto integrate = [](std::array<double, 6> p, const std::array<double, 3>& av, double h, double dT) {
    p[0] += av[0] * dT;
    p[1] += av[1] * dT;
    p[2] = h;
    p[3] += av[2] * dT;
    p[4] = p[5] = 0.0;
    return p;
  };
  const auto mid = integrate(pose, avg, cmd[2], tMid);
  const auto fin = integrate(mid, {vgx, vgy, cmd[3]}, cmd[2], horizon - tMid);
  TargetTrajectories tt;
  tt.timeTrajectory = {initTime, initTime + tMid, initTime + horizon};
  for (const auto& p : {pose, mid, fin}) {
    vector_t x(m.nx, 0.0);
    for (int k = 0; k < 6; ++k) {
      x[k] = momentum[k];
      x[6 + k] = p[k];
    }
    for (int j = 0; j < nj; ++j) x[12 + j] = m.defaultJointState[j];
    tt.stateTrajectory.push_back(x);
  }
  return tt;
}
// Ab^-1 * x0[0..6) for a centroidal state, from the device flow map at zero input: flow(x0, 0)[6..12) = Ab^-1 (m hbar)
inline std::array<double, 6> centroidalBaseVelocity(const HostModel& m, const vector_t& x0, int device = 0) {
  std::array<double, 6> bv{0, 0, 0, 0, 0, 0};
  bool zero = true;
  for (int k = 0; k < 6; ++k) zero = zero && x0[k] == 0.0;
  if (zero) return bv;
  if (m.cen.model_type == 1) {
    // SingleRigidBodyDynamics: Ab = [[m 1, m [R r]x T], [0, R I R' T]] in closed form (ModelHelperFunctions.cpp:61-79)
    const double z = x0[9], y = x0[10], xr = x0[11];
    const double cz = std::cos(z), sz = std::sin(z), cy = std::cos(y), sy = std::sin(y), cx = std::cos(xr), sx = std::sin(xr);
    const double R[9] = {cz * cy, cz * sy * sx - sz * cx, cz * sy * cx + sz * sx, sz * cy, sz * sy * sx + cz * cx, sz * sy * cx - cz * sx, -sy, cy * sx, cy * cx};
    const double Sb[9] = {-sy, 0, 1, cy * sx, cx, 0, cy * cx, -sx, 0};
    auto mul = [](const double* A, const double* B, double* C) {
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
    };
    double T[9], RI[9], Rt[9], RIRt[9], A22[9];
    mul(R, Sb, T);
    mul(R, m.cen.inertia_nominal, RI);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) Rt[3 * i + j] = R[3 * j + i];
    mul(RI, Rt, RIRt);
    mul(RIRt, T, A22);
    const double* r0 = m.cen.com_to_base_nominal;
    const double rw[3] = {R[0] * r0[0] + R[1] * r0[1] + R[2] * r0[2], R[3] * r0[0] + R[4] * r0[1] + R[5] * r0[2], R[6] * r0[0] + R[7] * r0[1] + R[8] * r0[2]};
    // w = A22^-1 h_ang (cofactor inverse)
    const double* A = A22;
    const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
    const double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
    const double inv[9] = {c00 / det, (A[2] * A[7] - A[1] * A[8]) / det, (A[1] * A[5] - A[2] * A[4]) / det,
                           c01 / det, (A[0] * A[8] - A[2] * A[6]) / det, (A[2] * A[3] - A[0] * A[5]) / det,
                           c02 / det, (A[1] * A[6] - A[0] * A[7]) / det, (A[0] * A[4] - A[1] * A[3]) / det};
    double w[3], Tw[3];
    for (int i = 0; i < 3; ++i) w[i] = inv[3 * i] * x0[3] + inv[3 * i + 1] * x0[4] + inv[3 * i + 2] * x0[5];
    for (int i = 0; i < 3; ++i) Tw[i] = T[3 * i] * w[0] + T[3 * i + 1] * w[1] + T[3 * i + 2] * w[2];
    const double cr[3] = {rw[1] * Tw[2] - rw[2] * Tw[1], rw[2] * Tw[0] - rw[0] * Tw[2], rw[0] * Tw[1] - rw[1] * Tw[0]};
    for (int i = 0; i < 3; ++i) {
      bv[i] = (x0[i] - m.totalMass * cr[i]) / m.totalMass;
      bv[3 + i] = w[i];
    }
    return bv;
  }
  vector_t u(m.nu, 0.0), xd(m.nx, 0.0);
  if (b200sqp_centroidal_flow_map(&m.desc, 1, x0.data(), u.data(), xd.data(), nullptr, nullptr, device) != 0)
    throw std::runtime_error(std::string("[b200sqp::host] centroidalBaseVelocity: ") + b200sqp_last_error());
  for (int k = 0; k < 6; ++k) bv[k] = xd[6 + k] / m.totalMass;
  return bv;
}

// WeightCompInitializer: weight-compensating normal forces on the stance feet
inline vector_t weightCompensatingInput(const HostModel& m, bool left, bool right) {
  vector_t u(m.nu, 0.0);
  const int ns = int(left) + int(right);
  if (ns > 0) {
    const double fz = m.totalMass * 9.81 / ns;
    if (left) u[2] = fz;
    if (right) u[8] = fz;
  }
  return u;
}

// LinearInterpolation::interpolate: zero-order extrapolation; for duplicated times the lower range ( ] is selected; tiny intervals snap
inline vector_t linearInterpolate(double t, const vector_t& times, const std::vector<vector_t>& data) {
  if (times.size() <= 1) return data.front();
  const int index = static_cast<int>(lowerBoundIndex(times, t)) - 1;  // lookup::findIntervalInTimeArray
  const int last = static_cast<int>(times.size()) - 1;
  int idx;
  double alpha;
  if (index < 0) {
    idx = 0;
    alpha = 1.0;
  } else if (index < last) {
    const double length = times[index + 1] - times[index], tillNext = times[index + 1] - t;
    idx = index;
    alpha = (length > 2.0 * kWeakEps) ? tillNext / length : (tillNext < 0.5 * length ? 0.0 : 1.0);
  } else {
    idx = std::max(last - 1, 0);
    alpha = 0.0;
  }
  vector_t out(data[idx].size());
  for (size_t k = 0; k < out.size(); ++k) out[k] = alpha * data[idx][k] + (1.0 - alpha) * data[idx + 1][k];
  return out;
}

struct PrimalSolution {  // ocs2_oc/include/ocs2_oc/oc_data/PrimalSolution.h:43-106 (feed-forward controller data only)
  vector_t timeTrajectory_;
  std::vector<vector_t> stateTrajectory_, inputTrajectory_;
  std::vector<uint8_t> postEventIndices_;  // node event annotations
  ModeSchedule modeSchedule_;
  void clear() { *this = PrimalSolution(); }
};

// ocs2::TrajectorySpreading (ocs2_oc/src/trajectory_adjustment/TrajectorySpreading.cpp:49-166,268-369; templates in
// include/ocs2_oc/trajectory_adjustment/TrajectorySpreading.h:124-183): adapts trajectories computed for an old mode schedule to a new one
// by spreading the values next to the moved event times and truncating where the mode sequences stop matching.
class TrajectorySpreading {
 public:
  struct Status {
    bool willTruncate = false, willPerformTrajectorySpreading = false;
  };
  Status set(const ModeSchedule& oldMs, const ModeSchedule& newMs, const vector_t& oldTime) {
    const double t0 = oldTime.front(), tf = oldTime.back();
    const int oldFirst = static_cast<int>(upperBoundIndex(oldMs.eventTimes, t0)), oldLast = static_cast<int>(upperBoundIndex(oldMs.eventTimes, tf));
    const int newFirst = static_cast<int>(upperBoundIndex(newMs.eventTimes, t0)), newLast = static_cast<int>(upperBoundIndex(newMs.eventTimes, tf));
    int oldStart = oldFirst;
    const int newStart = newFirst;
    int w = 0;
    while (oldStart < static_cast<int>(oldMs.modeSequence.size())) {
      w = 0;   // std::mismatch over [oldStart, oldLast] x [newStart, newLast]
      while (oldStart + w <= oldLast && newStart + w <= newLast && oldMs.modeSequence[oldStart + w] == newMs.modeSequence[newStart + w]) ++w;
      if (w > 0) break;
      ++oldStart;
    }
    vector_t oldM, newM;
    if (w > 0) {
      oldM.assign(oldMs.eventTimes.begin() + oldStart, oldMs.eventTimes.begin() + oldStart + w - 1);
      newM.assign(newMs.eventTimes.begin() + newStart, newMs.eventTimes.begin() + newStart + w - 1);
    }
    if (w > 0 && oldStart > oldFirst) {
      oldM.insert(oldM.begin(), oldMs.eventTimes[oldStart - 1]);
      newM.insert(newM.begin(), t0 - 1e-4);
    }
    const bool oldLastMatched = (oldStart + w - 1 == oldLast), newLastMatched = (newStart + w - 1 == newLast);
    // (w == 0 erases the whole trajectory; the reference indexes eventTimes out of range there, so that case is skipped)
    if (w > 0 && !oldLastMatched && (newLastMatched || oldMs.eventTimes[oldStart + w - 1] < newMs.eventTimes[newStart + w - 1])) {
      oldM.push_back(oldMs.eventTimes[oldStart + w - 1]);
      newM.push_back(newLastMatched ? tf + 1e-4 : newMs.eventTimes[newStart + w - 1]);
    }
    eraseFrom_ = oldTime.size();
    if (w == 0) eraseFrom_ = 0;
    else if (!newLastMatched) eraseFrom_ = lowerBoundIndex(oldTime, newMs.eventTimes[newStart + w - 1]);
    computeSpreadingStrategy(oldTime, oldM, newM);
    status_.willTruncate = eraseFrom_ < oldTime.size();
    status_.willPerformTrajectorySpreading = !valueIdx_.empty();
    return status_;
  }
  template <class T>
  void adjustTrajectory(std::vector<T>& traj) const {
    traj.erase(traj.begin() + eraseFrom_, traj.end());
    std::vector<T> values;
    for (size_t i : valueIdx_) values.push_back(traj[i]);
    for (size_t i = 0; i < valueIdx_.size(); ++i)
      for (size_t j = begin_[i]; j < end_[i]; ++j) traj[j] = values[i];
  }
  void adjustTimeTrajectory(vector_t& time) const {
    time.erase(time.begin() + eraseFrom_, time.end());
    for (size_t i = 0; i < postEventIndices_.size(); ++i) {
      time[postEventIndices_[i] - 1] = matchedEventTimes_[i];
      time[postEventIndices_[i]] = std::min(matchedEventTimes_[i] + kWeakEps, time.back());
    }
  }
  const std::vector<size_t>& getPostEventIndices() const { return postEventIndices_; }

 private:
  static std::vector<size_t> findPostEventIndices(const vector_t& eventTimes, const vector_t& time) {
    std::vector<size_t> out(eventTimes.size());
    for (size_t i = 0; i < eventTimes.size(); ++i)
      out[i] = (i == eventTimes.size() - 1 && eventTimes[i] == time.back()) ? time.size() - 1 : upperBoundIndex(time, eventTimes[i]);
    return out;
  }
  void computeSpreadingStrategy(const vector_t& oldTime, const vector_t& oldM, const vector_t& newM) {
    begin_.clear();
    end_.clear();
    valueIdx_.clear();
    postEventIndices_.clear();
    matchedEventTimes_.clear();
    const auto oldPost = findPostEventIndices(oldM, oldTime), newPost = findPostEventIndices(newM, oldTime);
    for (size_t j = 0; j < oldPost.size(); ++j) {
      if (newPost[j] < oldPost[j]) {          // backward spreading
        begin_.push_back(newPost[j]);
        end_.push_back(std::min(oldPost[j], eraseFrom_));
        valueIdx_.push_back(oldPost[j]);
      } else if (newPost[j] > oldPost[j]) {   // forward spreading
        begin_.push_back(j == 0 ? oldPost[j] : std::max(oldPost[j], newPost[j - 1]));
        end_.push_back(newPost[j]);
        valueIdx_.push_back(oldPost[j] - 1);
      }
      if (newPost[j] != 0 && newPost[j] < eraseFrom_) {
        postEventIndices_.push_back(newPost[j]);
        matchedEventTimes_.push_back(newM[j]);
      }
    }
  }
  Status status_;
  size_t eraseFrom_ = 0;
  std::vector<size_t> begin_, end_, valueIdx_, postEventIndices_;
  vector_t matchedEventTimes_;
};

// trajectorySpread(oldModeSchedule, newModeSchedule, primalSolution) (TrajectorySpreadingHelperFunctions.h:124-143), called at the top of
// SqpSolver::runImpl (SqpSolver.cpp:211-213).  Note the reference's convention clash, reproduced: the SQP's primal solution stores the
// pre- and post-event samples at the SAME time (toPrimalSolution -> toTime), so the "post-event index" found by upper_bound is the sample
// AFTER the post-event node and adjustTimeTrajectory moves that sample's time to event + eps, even when the schedules are identical.
inline TrajectorySpreading::Status trajectorySpread(const ModeSchedule& oldMs, const ModeSchedule& newMs, PrimalSolution& primal,
                                                   std::vector<size_t>* postEventIndices = nullptr) {
  TrajectorySpreading ts;
  const auto status = ts.set(oldMs, newMs, primal.timeTrajectory_);
  primal.modeSchedule_ = newMs;
  ts.adjustTrajectory(primal.stateTrajectory_);
  ts.adjustTrajectory(primal.inputTrajectory_);
  ts.adjustTimeTrajectory(primal.timeTrajectory_);
  if (postEventIndices) *postEventIndices = ts.getPostEventIndices();
  return status;
}

// One MPC instance in the layout of b200sqp_upload_instances
struct Instance {
  vector_t x0, x_init, u_init, t_nodes, swing_ref, impact_factor, arm_phase, x_ref;
  std::vector<uint8_t> node_event, contact_flags;
  ModeSchedule modeSchedule;
  int n_nodes() const { return static_cast<int>(t_nodes.size()); }
};

// multiple_shooting::initializeStateInputTrajectories: interpolate the previous primal solution where it overlaps the new horizon,
// WeightCompInitializer for the tail (and for everything on a cold start)
inline void initializeStateInputTrajectories(const HostModel& m, const vector_t& x0, const std::vector<AnnotatedTime>& td,
                                             const std::vector<uint8_t>& contact, const PrimalSolution* previous, vector_t& xs, vector_t& us) {
  const int n = static_cast<int>(td.size()), nx = m.nx, nu = m.nu;
  xs.assign(static_cast<size_t>(n) * nx, 0.0);
  us.assign(static_cast<size_t>(n - 1) * nu, 0.0);
  double tStateTill = td[0].time, tInputTill = td[0].time;
  const bool warm = previous && previous->timeTrajectory_.size() >= 2;
  if (warm) {
    tStateTill = previous->timeTrajectory_.back();
    tInputTill = previous->timeTrajectory_[previous->timeTrajectory_.size() - 2];
  }
  const double tInit = getIntervalStart(td[0]);
  vector_t xcur = (tInit < tStateTill) ? linearInterpolate(tInit, previous->timeTrajectory_, previous->stateTrajectory_) : x0;
  std::copy(xcur.begin(), xcur.end(), xs.begin());
  for (int i = 0; i < n - 1; ++i) {
    vector_t u(nu, 0.0);
    if (td[i].event != EV_PRE) {
      const double t = getIntervalStart(td[i]), tNext = getIntervalEnd(td[i + 1]);
      if (t > tInputTill || tNext > tStateTill) {
        u = weightCompensatingInput(m, contact[2 * i], contact[2 * i + 1]);
      } else {
        u = linearInterpolate(t, previous->timeTrajectory_, previous->inputTrajectory_);
        xcur = linearInterpolate(tNext, previous->timeTrajectory_, previous->stateTrajectory_);
      }
    }
    std::copy(u.begin(), u.end(), us.begin() + static_cast<size_t>(i) * nu);
    std::copy(xcur.begin(), xcur.end(), xs.begin() + static_cast<size_t>(i + 1) * nx);
  }
}

// multiple_shooting::toPrimalSolution: inputs at PreEvent nodes repeat the previous input, the last input is repeated
inline PrimalSolution toPrimalSolution(const Instance& inst, const double* x, const double* u, int nx, int nu) {
  PrimalSolution p;
  const int n = inst.n_nodes();
  p.timeTrajectory_ = inst.t_nodes;
  p.postEventIndices_ = inst.node_event;
  p.modeSchedule_ = inst.modeSchedule;
  for (int i = 0; i < n; ++i) p.stateTrajectory_.emplace_back(x + static_cast<size_t>(i) * nx, x + static_cast<size_t>(i + 1) * nx);
  for (int i = 0; i < n - 1; ++i) {
    if (inst.node_event[i] == EV_PRE && i > 0) p.inputTrajectory_.push_back(p.inputTrajectory_.back());
    else p.inputTrajectory_.emplace_back(u + static_cast<size_t>(i) * nu, u + static_cast<size_t>(i + 1) * nu);
  }
  p.inputTrajectory_.push_back(p.inputTrajectory_.back());
  return p;
}

// The per-instance reference manager: gait schedule + swing planner + target trajectories (SwitchedModelReferenceManager)
class SwitchedModelReferenceManager {
 public:
  explicit SwitchedModelReferenceManager(const HostModel& m) : m_(&m), planner_(m.swing) {}
  // GaitReceiver / insertModeSequenceTemplate: schedule `gait` from startTime on (tiled out to finalTime)
  void setGait(const std::string& gait, double startTime, double finalTime) {
    if (gait == "stance") return;
    auto it = m_->gaits.find(gait);
    if (it == m_->gaits.end()) throw std::invalid_argument("[b200sqp::host] unknown gait '" + gait + "'");
    gaitSchedule_.insertModeSequenceTemplate(it->second, startTime, finalTime);
  }
  void setTargetTrajectories(TargetTrajectories tt) { targets_ = std::move(tt); }
  const TargetTrajectories& getTargetTrajectories() const { return targets_; }
  // ReferenceManager::preSolverRun -> modifyReferences: mode schedule over [t0 - T, tf + T], swing plan for it
  void preSolverRun(double initTime, double finalTime) {
    const double T = finalTime - initTime;
    modeSchedule_ = gaitSchedule_.getModeSchedule(initTime - T, finalTime + T);
    planner_.update(modeSchedule_, 0.0);
  }
  const ModeSchedule& getModeSchedule() const { return modeSchedule_; }
  const SwingTrajectoryPlanner& getSwingTrajectoryPlanner() const { return planner_; }

 private:
  const HostModel* m_;
  GaitSchedule gaitSchedule_;
  SwingTrajectoryPlanner planner_;
  ModeSchedule modeSchedule_;
  TargetTrajectories targets_;
};

// everything runImpl needs for one instance (time grid with event nodes, node data, initial guess)
inline Instance buildInstance(const HostModel& m, const SwitchedModelReferenceManager& rm, double initTime, const vector_t& x0, double finalTime,
                              double dt, const PrimalSolution* previous) {
  Instance I;
  I.modeSchedule = rm.getModeSchedule();
  const auto td = timeDiscretizationWithEvents(initTime, finalTime, dt, I.modeSchedule.eventTimes);
  const int n = static_cast<int>(td.size()), nx = m.nx;
  I.x0 = x0;
  I.t_nodes.resize(n);
  I.node_event.resize(n);
  I.contact_flags.resize(2 * n);
  I.swing_ref.resize(6 * n);
  I.impact_factor.resize(2 * n);
  I.arm_phase.resize(n);
  I.x_ref.resize(static_cast<size_t>(n) * nx);
  const auto& planner = rm.getSwingTrajectoryPlanner();
  for (int i = 0; i < n; ++i) {
    I.t_nodes[i] = td[i].time;
    I.node_event[i] = td[i].event;
    const double t = getIntervalStart(td[i]);
    const auto c = modeNumber2StanceLeg(I.modeSchedule.modeAtTime(t));
    for (int leg = 0; leg < 2; ++leg) {
      I.contact_flags[2 * i + leg] = c[leg];
      const auto z = planner.zReference(leg, t);
      for (int k = 0; k < 3; ++k) I.swing_ref[(2 * i + leg) * 3 + k] = z[k];
      I.impact_factor[2 * i + leg] = planner.impactProximityFactor(leg, t);
    }
    I.arm_phase[i] = std::sin(2.0 * M_PI * (getPhaseVariable(I.modeSchedule, t) - 0.15));
    const vector_t xr = rm.getTargetTrajectories().getDesiredState(t);
    std::copy(xr.begin(), xr.end(), I.x_ref.begin() + static_cast<size_t>(i) * nx);
  }
  initializeStateInputTrajectories(m, x0, td, I.contact_flags, previous, I.x_init, I.u_init);
  return I;
}

}  // namespace b200sqp::host
