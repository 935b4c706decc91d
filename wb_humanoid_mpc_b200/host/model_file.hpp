// Host layer (C++17, no dependencies): the flat model file -> b200sqp_model_desc + the reference-manager parameters.
//
// The reference builds its OptimalControlProblem from URDF + task.info + reference.info + gait.info inside WBMpcInterface
// (humanoid_nmpc/humanoid_wb_mpc/src/WBMpcInterface.cpp:60-199).  A host linked against ocs2 would fill b200sqp_model_desc from those
// objects; without Pinocchio/boost in this image the same data is read from the flat text file that
// tools/make_model_data.py derives from the reference's config files (wb_humanoid_mpc_b200/data/g1_wb_model.txt).
#pragma once
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/b200sqp.h"

namespace b200sqp::host {

struct GaitTemplate {  // ModeSequenceTemplate (humanoid_common_mpc/include/humanoid_common_mpc/gait/ModeSequenceTemplate.h)
  std::vector<int> modeSequence;
  std::vector<double> switchingTimes;
};

struct SwingTrajectoryConfig {  // SwingTrajectoryPlanner::Config (task.info swing_trajectory_config)
  double liftOffVelocity, touchDownVelocity, swingHeight, touchDownHeightOffset, swingTimeScale;
  double ipfLiftOffVelocity, ipfTouchDownVelocity, ipfMidPointValue;
};

enum Mode { FLY = 0, RF = 1, LF = 2, STANCE = 3 };  // humanoid_common_mpc/gait/MotionPhaseDefinition.h
inline int modeFromString(const std::string& s) {
  if (s == "FLY") return FLY;
  if (s == "RF") return RF;
  if (s == "LF") return LF;
  if (s == "STANCE") return STANCE;
  throw std::invalid_argument("unknown mode name '" + s + "'");
}

struct HostModel {
  std::string name;
  int nj = 0, nx = 0, nu = 0;
  b200sqp_model_desc desc{};
  bool centroidal = false;    // kind 1: the centroidal MPC (nx = nu = 12 + nj); `cen` is then the second argument of b200sqp_cen_create
  b200sqp_cen_desc cen{};
  double totalMass = 0.0;
  std::vector<double> initialState, defaultJointState;
  double defaultBaseHeight = 0.0;
  SwingTrajectoryConfig swing{};
  double dt = 0.0, timeHorizon = 0.0;
  b200sqp_settings sqpSettings{};
  std::map<std::string, GaitTemplate> gaits;
};

// throws std::invalid_argument like the reference interfaces do on missing files (WBMpcInterface.cpp:72-91)
inline HostModel loadModelFile(const std::string& path) {
  std::ifstream in(path);
  if (!in) throw std::invalid_argument("[b200sqp::host] model file not found: " + path);
  std::map<std::string, std::vector<double>> rec;
  HostModel m;
  std::string line;
  while (std::getline(in, line)) {
    std::istringstream ss(line);
    std::string key;
    if (!(ss >> key)) continue;
    if (key == "name") {
      int one;
      ss >> one >> m.name;
    } else if (key == "gait") {
      std::string gname;
      size_t n;
      ss >> gname >> n;
      GaitTemplate g;
      for (size_t i = 0; i < n; ++i) {
        std::string mode;
        ss >> mode;
        g.modeSequence.push_back(modeFromString(mode));
      }
      g.switchingTimes.resize(n + 1);
      for (auto& t : g.switchingTimes) ss >> t;
      if (!ss) throw std::invalid_argument("[b200sqp::host] malformed gait record '" + gname + "'");
      m.gaits[gname] = g;
    } else {
      size_t n;
      ss >> n;
      std::vector<double> v(n);
      for (auto& x : v) ss >> x;
      if (!ss) throw std::invalid_argument("[b200sqp::host] malformed record '" + key + "'");
      rec[key] = v;
    }
  }
  auto get = [&](const char* k, size_t n = 0) -> const std::vector<double>& {
    auto it = rec.find(k);
    if (it == rec.end()) throw std::invalid_argument(std::string("[b200sqp::host] model file lacks '") + k + "'");
    if (n && it->second.size() != n) throw std::invalid_argument(std::string("[b200sqp::host] record '") + k + "' has the wrong length");
    return it->second;
  };
  m.nj = static_cast<int>(get("nj", 1)[0]);
  m.nx = static_cast<int>(get("nx", 1)[0]);
  m.nu = static_cast<int>(get("nu", 1)[0]);
  const int nb = m.nj + 1;
  if (nb > 32) throw std::invalid_argument("[b200sqp::host] too many bodies");
  b200sqp_model_desc& d = m.desc;
  std::memset(&d, 0, sizeof(d));
  d.nj = m.nj;
  const auto &par = get("parent", nb), &jR = get("joint_R", 9 * nb), &jp = get("joint_p", 3 * nb), &ax = get("joint_axis", 3 * nb),
             &ms = get("mass", nb), &com = get("com", 3 * nb), &I = get("inertia", 9 * nb);
  for (int i = 0; i < nb; ++i) {
    d.parent[i] = static_cast<int32_t>(par[i]);
    for (int k = 0; k < 9; ++k) {
      d.joint_R[i][k] = jR[9 * i + k];
      d.inertia[i][k] = I[9 * i + k];
    }
    for (int k = 0; k < 3; ++k) {
      d.joint_p[i][k] = jp[3 * i + k];
      d.joint_axis[i][k] = ax[3 * i + k];
      d.com[i][k] = com[3 * i + k];
    }
    d.mass[i] = ms[i];
    m.totalMass += ms[i];
  }
  const auto &ql = get("q_lower", m.nj), &qu = get("q_upper", m.nj);
  for (int j = 0; j < m.nj; ++j) {
    d.q_lower[j] = ql[j];
    d.q_upper[j] = qu[j];
  }
  const auto& fb = get("frame_body");
  const auto& fp = get("frame_p", 3 * fb.size());
  if (fb.size() > 16) throw std::invalid_argument("[b200sqp::host] too many frames");
  d.n_frames = static_cast<int32_t>(fb.size());
  for (size_t f = 0; f < fb.size(); ++f) {
    d.frame_body[f] = static_cast<int32_t>(fb[f]);
    for (int k = 0; k < 3; ++k) d.frame_p[f][k] = fp[3 * f + k];
  }
  d.gravity = get("gravity", 1)[0];
  for (int k = 0; k < 4; ++k) d.contact_rect[k] = get("contact_rect", 4)[k];
  for (int i = 0; i < m.nx; ++i) {
    d.Q_diag[i] = get("Q_diag", m.nx)[i];
    d.Qf_diag[i] = get("Qf_diag", m.nx)[i];
  }
  for (int i = 0; i < m.nu; ++i) d.R_diag[i] = get("R_diag", m.nu)[i];
  const auto& g = get("foot_gains", 8);
  d.foot_gain_pos_z = g[0];
  d.foot_gain_ori = g[1];
  d.foot_gain_linvel_z = g[2];
  d.foot_gain_linvel_xy = g[3];
  d.foot_gain_angvel = g[4];
  d.foot_gain_linacc_z = g[5];
  d.foot_gain_linacc_xy = g[6];
  d.foot_gain_angacc = g[7];
  for (int i = 0; i < 18; ++i) d.foot_cost_w[i] = get("foot_cost_weights", 18)[i];
  const auto& fr = get("friction", 5);
  d.fric_coeff = fr[0];
  d.fric_mu = fr[1];
  d.fric_delta = fr[2];
  d.fric_reg = fr[3];
  d.fric_hess_shift = fr[4];
  d.momxy_mu = get("moment_xy", 2)[0];
  d.momxy_delta = get("moment_xy", 2)[1];
  d.jlim_mu = get("joint_limits", 2)[0];
  d.jlim_delta = get("joint_limits", 2)[1];
  const auto& c = get("collision", 4);
  d.coll_mu = c[0];
  d.coll_delta = c[1];
  d.coll_r_foot = c[2];
  d.coll_r_knee = c[3];
  for (int i = 0; i < 4; ++i) d.arm_swing_joint[i] = static_cast<int32_t>(get("arm_swing_joints", 4)[i]);
  m.centroidal = rec.count("kind") && rec["kind"].size() == 1 && rec["kind"][0] == 1.0;
  if (m.centroidal) {
    b200sqp_cen_desc& c2 = m.cen;
    std::memset(&c2, 0, sizeof(c2));
    c2.torso_frame = static_cast<int32_t>(get("cen_torso_frame", 1)[0]);
    for (int k = 0; k < 9; ++k) c2.torso_R[k] = get("cen_torso_R", 9)[k];
    for (int k = 0; k < 12; ++k) c2.torso_w[k] = get("cen_torso_w", 12)[k];
    c2.icp_weight = get("cen_icp_weight", 1)[0];
    for (int s2 = 0; s2 < 2; ++s2)
      for (int k = 0; k < 6; ++k) {
        c2.torque_joint[s2][k] = static_cast<int32_t>(get("cen_torque_joint", 12)[6 * s2 + k]);
        c2.torque_w[s2][k] = get("cen_torque_w", 12)[6 * s2 + k];
      }
    c2.model_type = static_cast<int32_t>(get("cen_model_type", 1)[0]);
    for (int k = 0; k < 9; ++k) c2.inertia_nominal[k] = get("cen_inertia_nominal", 9)[k];
    for (int k = 0; k < 3; ++k) c2.com_to_base_nominal[k] = get("cen_com_to_base_nominal", 3)[k];
  }
  m.initialState = get("x_init", m.nx);
  m.defaultJointState = get("default_joint_state", m.nj);
  m.defaultBaseHeight = get("default_base_height", 1)[0];
  const auto& sw = get("swing", 8);
  m.swing = SwingTrajectoryConfig{sw[0], sw[1], sw[2], sw[3], sw[4], sw[5], sw[6], sw[7]};
  const auto& sq = get("sqp", 6);
  m.dt = sq[0];
  m.timeHorizon = sq[5];
  b200sqp_default_settings(&m.sqpSettings);   // sqp::Settings defaults, then task.info:77-94
  m.sqpSettings.sqp_iteration = static_cast<int32_t>(sq[1]);
  m.sqpSettings.delta_tol = sq[2];
  m.sqpSettings.g_max = sq[3];
  m.sqpSettings.g_min = sq[4];
  return m;
}

}  // namespace b200sqp::host
