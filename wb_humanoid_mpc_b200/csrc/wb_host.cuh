// Host-side conversion of the public model description into the device constant block.
#pragma once
#include <cmath>
#include <cstring>

#include "wb_model.cuh"

namespace b200sqp {

// returns nullptr on success, else a static error string
inline const char* makeDeviceModel(const b200sqp_model_desc& d, WbDeviceModel& m) {
  if (d.nj != NJ) return "this build supports nj = 23 (Unitree G1 whole-body MPC model)";
  if (d.n_frames != NFRAMES) return "expected 10 operational frames (2 x {contact, p1, p2}, 2 ankles, 2 knees)";
  std::memset(&m, 0, sizeof(m));
  for (int i = 0; i < NB; ++i) {
    m.parent[i] = d.parent[i];
    if (i > 0 && (d.parent[i] < 0 || d.parent[i] >= i)) return "bodies must be ordered parent-before-child";
    for (int k = 0; k < 9; ++k) {
      m.jR[i][k] = d.joint_R[i][k];
      m.Icom[i][k] = d.inertia[i][k];
    }
    for (int k = 0; k < 3; ++k) {
      m.jp[i][k] = d.joint_p[i][k];
      m.axis[i][k] = d.joint_axis[i][k];
      m.com[i][k] = d.com[i][k];
    }
    m.mass[i] = d.mass[i];
    m.mtot += d.mass[i];
  }
  for (int i = 0; i < NB; ++i) {
    unsigned mask = 0;
    for (int j = i; j < NB; ++j) {
      int a = j;
      while (a > i) a = m.parent[a];
      if (a == i) mask |= 1u << j;
    }
    m.subtree[i] = mask;
    int tmp[NB], len = 0;
    for (int a = i; a > 0; a = m.parent[a]) tmp[len++] = a;
    if (len > 8) return "kinematic chains deeper than 8 joints are not supported";
    m.pathLen[i] = len;
    for (int t = 0; t < len; ++t) m.path[i][t] = tmp[len - 1 - t];
  }
  // kinematic layout assumed by the chain kernels: legs 1-6 / 7-12, waist 13-15, arms 16-19 / 20-23
  const int expect[NB] = {-1, 0, 1, 2, 3, 4, 5, 0, 7, 8, 9, 10, 11, 0, 13, 14, 15, 16, 17, 18, 15, 20, 21, 22};
  for (int i = 0; i < NB; ++i)
    if (m.parent[i] != expect[i]) return "kinematic tree layout differs from the G1 leg/leg/waist/arm/arm layout this build is specialised for";
  m.gravity = d.gravity;
  for (int j = 0; j < NJ; ++j) {
    m.qlo[j] = d.q_lower[j];
    m.qhi[j] = d.q_upper[j];
  }
  for (int f = 0; f < NFRAMES; ++f) {
    if (d.frame_body[f] < 0 || d.frame_body[f] >= NB) return "frame_body[] entry is not a body index";
    m.frameBody[f] = d.frame_body[f];
    for (int k = 0; k < 3; ++k) m.frameP[f][k] = d.frame_p[f][k];
  }
  for (int c = 0; c < 2; ++c) {
    int b = m.frameBody[3 * c];
    for (int l = LEG_LEN - 1; l >= 0; --l) {
      m.legBody[c][l] = b;
      b = m.parent[b];
    }
    if (b != 0) return "contact frames must hang off a 6-joint leg attached to the base";
    for (int l = 0; l < LEG_LEN; ++l)
      if (m.legBody[c][l] != 1 + LEG_LEN * c + l) return "the leg joints must be bodies 1..12 (left leg first): the structured penalty rows assume it";
  }
  for (int k = 0; k < 4; ++k) m.rect[k] = d.contact_rect[k];
  for (int i = 0; i < NX; ++i) {
    m.Qd[i] = d.Q_diag[i];
    m.Qfd[i] = d.Qf_diag[i];
  }
  for (int i = 0; i < NU; ++i) m.Rd[i] = d.R_diag[i];
  m.gPosZ = d.foot_gain_pos_z;
  m.gOri = d.foot_gain_ori;
  m.gLinVelZ = d.foot_gain_linvel_z;
  m.gLinVelXY = d.foot_gain_linvel_xy;
  m.gAngVel = d.foot_gain_angvel;
  m.gLinAccZ = d.foot_gain_linacc_z;
  m.gLinAccXY = d.foot_gain_linacc_xy;
  m.gAngAcc = d.foot_gain_angacc;
  for (int k = 0; k < 18; ++k) m.footSqrtW[k] = std::sqrt(d.foot_cost_w[k]);
  m.fricCoeff = d.fric_coeff;
  m.fricMu = d.fric_mu;
  m.fricDelta = d.fric_delta;
  m.fricReg = d.fric_reg;
  m.fricShift = d.fric_hess_shift;
  m.momMu = d.momxy_mu;
  m.momDelta = d.momxy_delta;
  m.jlMu = d.jlim_mu;
  m.jlDelta = d.jlim_delta;
  m.collMu = d.coll_mu;
  m.collDelta = d.coll_delta;
  m.rFoot = d.coll_r_foot;
  m.rKnee = d.coll_r_knee;
  for (int k = 0; k < 4; ++k) {
    if (d.arm_swing_joint[k] < 0 || d.arm_swing_joint[k] >= NJ) return "arm_swing_joint[] entry is not a joint index";
    m.armJoint[k] = d.arm_swing_joint[k];
  }
  return nullptr;
}

}  // namespace b200sqp
