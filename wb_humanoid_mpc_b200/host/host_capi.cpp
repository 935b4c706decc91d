// C entry points of the C++ host layer (libb200sqp_host.so) so that tests/ and bench.py can drive b200sqp::host::SqpSolver and the instance
// builder through ctypes.  A C++ application includes SqpSolver.hpp directly; nothing here adds behaviour.
#include <cstring>
#include <string>

#include "model_from_config.hpp"
#include "MpcPolicyMsg.hpp"
#include "SqpLogging.hpp"
#include "SqpSolver.hpp"

using namespace b200sqp::host;

namespace {
thread_local std::string g_err;
template <class F>
int guarded(F f) {
  try {
    return f();
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
}  // namespace

extern "C" {

const char* b200host_last_error() { return g_err.c_str(); }

void* b200host_model_load(const char* path) {
  try {
    return new HostModel(loadModelFile(path));
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
// the whole-body (centroidal = 0) or centroidal (1) model straight from the reference's own config files (URDF + task.info + reference.info +
// gait.info; gait may be "")
void* b200host_model_from_config(const char* urdf, const char* task, const char* reference, const char* gait, int centroidal) {
  try {
    return new HostModel(loadModelFromConfig(urdf, task, reference, gait ? gait : "", centroidal != 0));
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
// everything of a HostModel that b200host_model_desc / _dims do not return, as one array (test hook): total mass, default base height, swing
// config (8), dt, horizon, initial state (nx), default joint state (nj), then per gait in name order: count, modes, switching times
int b200host_model_dump(void* m, double* out, int cap) {
  const HostModel& M = *static_cast<HostModel*>(m);
  std::vector<double> v{M.totalMass, M.defaultBaseHeight, M.swing.liftOffVelocity, M.swing.touchDownVelocity, M.swing.swingHeight,
                        M.swing.touchDownHeightOffset, M.swing.swingTimeScale, M.swing.ipfLiftOffVelocity, M.swing.ipfTouchDownVelocity,
                        M.swing.ipfMidPointValue, M.dt, M.timeHorizon};
  v.insert(v.end(), M.initialState.begin(), M.initialState.end());
  v.insert(v.end(), M.defaultJointState.begin(), M.defaultJointState.end());
  for (const auto& g : M.gaits) {
    v.push_back(static_cast<double>(g.second.modeSequence.size()));
    for (int mode : g.second.modeSequence) v.push_back(mode);
    v.insert(v.end(), g.second.switchingTimes.begin(), g.second.switchingTimes.end());
  }
  const int n = static_cast<int>(v.size());
  for (int i = 0; i < n && i < cap; ++i) out[i] = v[i];
  return n;
}
void b200host_model_free(void* m) { delete static_cast<HostModel*>(m); }
int b200host_model_dims(void* m, int* nx, int* nu, double* dt, double* horizon) {
  const HostModel& M = *static_cast<HostModel*>(m);
  *nx = M.nx;
  *nu = M.nu;
  *dt = M.dt;
  *horizon = M.timeHorizon;
  return 0;
}
int b200host_model_desc(void* m, b200sqp_model_desc* desc, b200sqp_settings* st) {
  const HostModel& M = *static_cast<HostModel*>(m);
  if (desc) *desc = M.desc;
  if (st) *st = M.sqpSettings;
  return 0;
}

// 1 and *cen filled for a centroidal model file, 0 for a whole-body one
int b200host_model_cen_desc(void* m, b200sqp_cen_desc* cen) {
  const HostModel& M = *static_cast<HostModel*>(m);
  if (M.centroidal && cen) *cen = M.cen;
  return M.centroidal ? 1 : 0;
}

// One instance, host only (no GPU): the arrays of b200sqp_upload_instances.  prev_* = the previous PrimalSolution (equal-length time, state
// and input trajectories) or prev_n = 0 for a cold start.  Returns the node count, or -1 (b200host_last_error).
namespace {
// target trajectories of either MPC; base_vel (6, may be null = zeros) is only read for a centroidal model
TargetTrajectories targetsFor(const HostModel& M, double t0, const vector_t& x0, const std::array<double, 4>& c, double horizon, const double* base_vel) {
  if (!M.centroidal) return commandedVelocityToTargetTrajectories(M, t0, x0, c, horizon);
  std::array<double, 6> bv{0, 0, 0, 0, 0, 0};
  if (base_vel) std::copy(base_vel, base_vel + 6, bv.begin());
  return commandedVelocityToTargetTrajectoriesCentroidal(M, t0, x0, c, horizon, bv);
}
}  // namespace

int b200host_build_instance(void* model, double t0, const double* x0, double horizon, const char* gait, double gait_start, const double* cmd,
                            int prev_n, const double* prev_t, const double* prev_x, const double* prev_u, int max_nodes, double* t_nodes,
                            uint8_t* node_event, uint8_t* contact, double* swing, double* impact, double* arm, double* xref, double* x_init,
                            double* u_init, const double* base_vel) {
  return guarded([&] {
    const HostModel& M = *static_cast<HostModel*>(model);
    SwitchedModelReferenceManager rm(M);
    const double tf = t0 + horizon;
    rm.setGait(gait, gait_start, tf + horizon);
    const vector_t x0v(x0, x0 + M.nx);
    std::array<double, 4> c{0.0, 0.0, M.defaultBaseHeight, 0.0};
    if (cmd) c = {cmd[0], cmd[1], cmd[2], cmd[3]};
    rm.setTargetTrajectories(targetsFor(M, t0, x0v, c, horizon, base_vel));
    rm.preSolverRun(t0, tf);
    PrimalSolution prev;
    if (prev_n > 0) {
      prev.timeTrajectory_.assign(prev_t, prev_t + prev_n);
      for (int i = 0; i < prev_n; ++i) {
        prev.stateTrajectory_.emplace_back(prev_x + static_cast<size_t>(i) * M.nx, prev_x + static_cast<size_t>(i + 1) * M.nx);
        prev.inputTrajectory_.emplace_back(prev_u + static_cast<size_t>(i) * M.nu, prev_u + static_cast<size_t>(i + 1) * M.nu);
      }
    }
    const Instance I = buildInstance(M, rm, t0, x0v, tf, M.dt, prev_n > 0 ? &prev : nullptr);
    const int n = I.n_nodes();
    if (n > max_nodes) throw std::runtime_error("b200host_build_instance: max_nodes too small");
    std::copy(I.t_nodes.begin(), I.t_nodes.end(), t_nodes);
    std::copy(I.node_event.begin(), I.node_event.end(), node_event);
    std::copy(I.contact_flags.begin(), I.contact_flags.end(), contact);
    std::copy(I.swing_ref.begin(), I.swing_ref.end(), swing);
    std::copy(I.impact_factor.begin(), I.impact_factor.end(), impact);
    std::copy(I.arm_phase.begin(), I.arm_phase.end(), arm);
    std::copy(I.x_ref.begin(), I.x_ref.end(), xref);
    std::copy(I.x_init.begin(), I.x_init.end(), x_init);
    std::copy(I.u_init.begin(), I.u_init.end(), u_init);
    return n;
  });
}

// createMpcPolicyMsg on flat arrays (test entry point): t [n], event [n], x [n][nx], u [n-1][nu], K [n-1][nu*nx] or null -> data [n][stride] (floats),
// post-event indices; returns the per-sample data length, -1 on error; evaluates the packed policy at sample `probe` for state x_probe into u_probe
int b200host_policy_msg(int n, int nx, int nu, const double* t, const uint8_t* event, const double* x, const double* u, const double* K, float* data, int data_cap,
                        uint16_t* post, int* n_post, int probe, const double* x_probe, double* u_probe) {
  return guarded([&] {
    Instance inst;
    inst.t_nodes.assign(t, t + n);
    inst.node_event.assign(event, event + n);
    const PrimalSolution p = toPrimalSolution(inst, x, u, nx, nu);
    const MpcFlattenedController m = createMpcPolicyMsg(p, t[0], vector_t(x, x + nx), K, nx, nu);
    const int stride = static_cast<int>(m.data.front().size());
    if (stride * n > data_cap) throw std::runtime_error("policy_msg: data buffer too small");
    for (int k = 0; k < n; ++k) std::copy(m.data[k].begin(), m.data[k].end(), data + static_cast<size_t>(k) * stride);
    *n_post = static_cast<int>(m.postEventIndices.size());
    std::copy(m.postEventIndices.begin(), m.postEventIndices.end(), post);
    const auto up = evaluatePolicySample(m, static_cast<size_t>(probe), vector_t(x_probe, x_probe + nx));
    std::copy(up.begin(), up.end(), u_probe);
    return stride;
  });
}

// trajectorySpread on a primal solution given as flat arrays (in place); returns the new length or -1
int b200host_trajectory_spread(int n_old_ev, const double* old_ev, const int* old_modes, int n_new_ev, const double* new_ev, const int* new_modes, int n,
                               int nx, int nu, double* t, double* x, double* u, int* flags /* willTruncate, willSpread */) {
  return guarded([&] {
    ModeSchedule oldMs, newMs;
    oldMs.eventTimes.assign(old_ev, old_ev + n_old_ev);
    oldMs.modeSequence.assign(old_modes, old_modes + n_old_ev + 1);
    newMs.eventTimes.assign(new_ev, new_ev + n_new_ev);
    newMs.modeSequence.assign(new_modes, new_modes + n_new_ev + 1);
    PrimalSolution p;
    p.timeTrajectory_.assign(t, t + n);
    for (int i = 0; i < n; ++i) {
      p.stateTrajectory_.emplace_back(x + static_cast<size_t>(i) * nx, x + static_cast<size_t>(i + 1) * nx);
      p.inputTrajectory_.emplace_back(u + static_cast<size_t>(i) * nu, u + static_cast<size_t>(i + 1) * nu);
    }
    const auto st = trajectorySpread(oldMs, newMs, p);
    const int m = static_cast<int>(p.timeTrajectory_.size());
    std::copy(p.timeTrajectory_.begin(), p.timeTrajectory_.end(), t);
    for (int i = 0; i < m; ++i) {
      std::copy(p.stateTrajectory_[i].begin(), p.stateTrajectory_[i].end(), x + static_cast<size_t>(i) * nx);
      std::copy(p.inputTrajectory_[i].begin(), p.inputTrajectory_[i].end(), u + static_cast<size_t>(i) * nu);
    }
    flags[0] = st.willTruncate;
    flags[1] = st.willPerformTrajectorySpreading;
    return m;
  });
}

void* b200host_solver_create(void* model, const b200sqp_settings* st, int batch, int device, int threads) {
  try {
    return new SqpSolver(*static_cast<HostModel*>(model), *st, batch, device, threads);
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
void b200host_solver_destroy(void* s) { delete static_cast<SqpSolver*>(s); }
int b200host_solver_set_gait(void* s, int b, const char* gait, double start, double final_time) {
  return guarded([&] {
    static_cast<SqpSolver*>(s)->getReferenceManager(b).setGait(gait, start, final_time);
    return 0;
  });
}
int b200host_solver_set_command(void* s, void* model, int b, double t0, const double* x0, const double* cmd, double horizon, const double* base_vel) {
  return guarded([&] {
    const HostModel& M = *static_cast<HostModel*>(model);
    static_cast<SqpSolver*>(s)->getReferenceManager(b).setTargetTrajectories(
        targetsFor(M, t0, vector_t(x0, x0 + M.nx), {cmd[0], cmd[1], cmd[2], cmd[3]}, horizon, base_vel));
    return 0;
  });
}
// Ab^-1 * x0[0..6) of a centroidal state (device call; zeros without touching the device when the momentum is zero)
int b200host_centroidal_base_velocity(void* model, const double* x0, double* out6) {
  return guarded([&] {
    const HostModel& M = *static_cast<HostModel*>(model);
    const auto bv = centroidalBaseVelocity(M, vector_t(x0, x0 + M.nx));
    std::copy(bv.begin(), bv.end(), out6);
    return 0;
  });
}
int b200host_solver_set_trajectory_spread(void* s, int on) {
  static_cast<SqpSolver*>(s)->setTrajectorySpread(on != 0);
  return 0;
}
int b200host_solver_set_exclusive_solve(void* s, int on) {
  static_cast<SqpSolver*>(s)->setExclusiveSolve(on != 0);
  return 0;
}
int b200host_solver_reset(void* s) {
  return guarded([&] {
    static_cast<SqpSolver*>(s)->reset();
    return 0;
  });
}
int b200host_solver_run(void* s, void* model, double t0, const double* x0s, double tf) {
  return guarded([&] {
    SqpSolver& S = *static_cast<SqpSolver*>(s);
    const int nx = static_cast<HostModel*>(model)->nx;
    std::vector<vector_t> xs(S.batch());
    for (int b = 0; b < S.batch(); ++b) xs[b].assign(x0s + static_cast<size_t>(b) * nx, x0s + static_cast<size_t>(b + 1) * nx);
    S.run(t0, xs, tf);
    return 0;
  });
}
int b200host_solver_n_nodes(void* s, int b) {
  return guarded([&] { return static_cast<int>(static_cast<SqpSolver*>(s)->primalSolution(b).timeTrajectory_.size()); });
}
int b200host_solver_get_primal(void* s, int b, double* t, double* x, double* u) {
  return guarded([&] {
    const PrimalSolution& p = static_cast<SqpSolver*>(s)->primalSolution(b);
    const size_t n = p.timeTrajectory_.size();
    std::copy(p.timeTrajectory_.begin(), p.timeTrajectory_.end(), t);
    for (size_t i = 0; i < n; ++i) {
      std::copy(p.stateTrajectory_[i].begin(), p.stateTrajectory_[i].end(), x + i * p.stateTrajectory_[i].size());
      std::copy(p.inputTrajectory_[i].begin(), p.inputTrajectory_[i].end(), u + i * p.inputTrajectory_[i].size());
    }
    return static_cast<int>(n);
  });
}
int b200host_solver_get_log(void* s, int b, double* out /* [iters][12] */, int max_iter) {
  return guarded([&] {
    const auto& L = static_cast<SqpSolver*>(s)->getIterationsLog(b);
    int k = 0;
    for (const StepInfo& si : L) {
      if (k >= max_iter) break;
      double* o = out + 12 * k++;
      o[0] = si.baseline.merit;
      o[1] = si.baseline.dynamicsViolationSSE;
      o[2] = si.baseline.equalityConstraintsSSE;
      o[3] = si.performanceAfterStep.merit;
      o[4] = si.performanceAfterStep.dynamicsViolationSSE;
      o[5] = si.performanceAfterStep.equalityConstraintsSSE;
      o[6] = si.stepSize;
      o[7] = si.stepType;
      o[8] = si.dx_norm;
      o[9] = si.du_norm;
      o[10] = si.armijoDescentMetric;
      o[11] = si.convergence;
    }
    return k;
  });
}
// sqp::Logger CSV (header + one line per instance and iteration of the last run) into buf; returns the length needed (without the NUL)
int b200host_solver_write_log(void* s, double time, char* buf, int cap) {
  return guarded([&] {
    std::ostringstream os;
    os << logHeader();
    writeLog(os, *static_cast<SqpSolver*>(s), time);
    const std::string str = os.str();
    if (buf && cap > 0) {
      const size_t n = std::min(str.size(), static_cast<size_t>(cap - 1));
      std::memcpy(buf, str.data(), n);
      buf[n] = 0;
    }
    return static_cast<int>(str.size());
  });
}
int b200host_solver_benchmarks(void* s, double* ms) {
  const Benchmarks b = static_cast<SqpSolver*>(s)->getBenchmarks();
  ms[0] = b.linearQuadraticApproximation;
  ms[1] = b.solveQp;
  ms[2] = b.linesearch;
  ms[3] = b.projectionShareOfLq;
  ms[4] = b.hostPreRun;
  ms[5] = b.hostPack;
  ms[6] = b.upload;
  ms[7] = b.solve;
  ms[8] = b.download;
  ms[9] = b.hostUnpack;
  return 0;
}
int b200host_solver_value_function(void* s, int b, double t, const double* x, int nx, double* dfdxx, double* dfdx) {
  return guarded([&] {
    const ValueFunction v = static_cast<SqpSolver*>(s)->getValueFunction(b, t, vector_t(x, x + nx));
    std::copy(v.dfdxx.begin(), v.dfdxx.end(), dfdxx);
    std::copy(v.dfdx.begin(), v.dfdx.end(), dfdx);
    return 0;
  });
}

}  // extern "C"
