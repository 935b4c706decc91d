// Host-side conversion of (b200sqp_model_desc, b200sqp_cen_desc) into the centroidal device constant block.
#pragma once
#include <memory>

#include "cen_solver.cuh"
#include "wb_host.cuh"

namespace b200sqp {

inline void cenModelFromWb(const WbDeviceModel& wm, CenModel& cm) {
  for (int i = 0; i < NB; ++i) {
    cm.parent[i] = wm.parent[i];
    cm.mass[i] = wm.mass[i];
    for (int k = 0; k < 9; ++k) {
      cm.jR[i][k] = wm.jR[i][k];
      cm.Icom[i][k] = wm.Icom[i][k];
    }
    for (int k = 0; k < 3; ++k) {
      cm.jp[i][k] = wm.jp[i][k];
      cm.axis[i][k] = wm.axis[i][k];
      cm.com[i][k] = wm.com[i][k];
    }
  }
  cm.mtot = wm.mtot;
  cm.modelType = 0;
  for (int k = 0; k < 9; ++k) cm.inertiaNominal[k] = 0.0;
  for (int k = 0; k < 3; ++k) cm.comToBaseNominal[k] = 0.0;
  for (int c = 0; c < 2; ++c) {
    cm.contactBody[c] = wm.frameBody[3 * c];
    for (int k = 0; k < 3; ++k) cm.contactP[c][k] = wm.frameP[3 * c][k];
  }
}

// returns nullptr on success, else a static error string
inline const char* makeCenDeviceModel(const b200sqp_model_desc& d, const b200sqp_cen_desc& c, CenDevModel& out) {
  if (d.n_frames != CEN_NFRAMES) return "the centroidal model needs 11 operational frames (the 10 whole-body frames + the task-space link)";
  if (c.torso_frame != NFRAMES) return "torso_frame must be the last entry (index 10) of the frame table";
  b200sqp_model_desc base = d;
  base.n_frames = NFRAMES;
  const std::unique_ptr<WbDeviceModel> wmp(new WbDeviceModel);
  WbDeviceModel& wm = *wmp;
  if (const char* e = makeDeviceModel(base, wm)) return e;
  std::memset(&out, 0, sizeof(out));
  CenOcpModel& m = out.ocp;
  cenModelFromWb(wm, m.kin);
  for (int f = 0; f < CEN_NFRAMES; ++f) {
    m.frameBody[f] = d.frame_body[f];
    if (d.frame_body[f] < 0 || d.frame_body[f] >= NB) return "frame body out of range";
    for (int k = 0; k < 3; ++k) m.frameP[f][k] = d.frame_p[f][k];
  }
  if (c.model_type != 0 && c.model_type != 1) return "model_type must be 0 (FullCentroidalDynamics) or 1 (SingleRigidBodyDynamics)";
  m.kin.modelType = c.model_type;
  for (int k = 0; k < 9; ++k) m.kin.inertiaNominal[k] = c.inertia_nominal[k];
  for (int k = 0; k < 3; ++k) m.kin.comToBaseNominal[k] = c.com_to_base_nominal[k];
  m.torsoFrame = c.torso_frame;
  for (int k = 0; k < 9; ++k) m.torsoR[k] = c.torso_R[k];
  for (int k = 0; k < 12; ++k) {
    if (c.torso_w[k] < 0 || d.foot_cost_w[k] < 0) return "negative task-space weight";
    m.torsoSqrtW[k] = std::sqrt(c.torso_w[k]);
    m.footSqrtW[k] = std::sqrt(d.foot_cost_w[k]);
  }
  if (c.icp_weight < 0) return "negative ICP weight";
  m.icpSqrtW = std::sqrt(c.icp_weight);
  for (int s = 0; s < 2; ++s)
    for (int k = 0; k < 6; ++k) {
      if (c.torque_joint[s][k] < 0 || c.torque_joint[s][k] >= NJ || c.torque_w[s][k] < 0) return "external-torque cost: joint index or weight out of range";
      m.tqJoint[s][k] = c.torque_joint[s][k];
      m.tqSqrtW[s][k] = std::sqrt(c.torque_w[s][k]);
    }
  for (int i = 0; i < CNX; ++i) {
    m.Qd[i] = d.Q_diag[i];
    m.Qfd[i] = d.Qf_diag[i];
    out.QfdPad[i] = d.Qf_diag[i];
  }
  for (int i = 0; i < CNU; ++i) m.Rd[i] = d.R_diag[i];
  m.gPosZ = d.foot_gain_pos_z;
  m.gOri = d.foot_gain_ori;
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
    m.rect[k] = d.contact_rect[k];
    m.armJoint[k] = d.arm_swing_joint[k];
  }
  for (int j = 0; j < NJ; ++j) {
    m.qlo[j] = d.q_lower[j];
    m.qhi[j] = d.q_upper[j];
  }
  return nullptr;
}

}  // namespace b200sqp
