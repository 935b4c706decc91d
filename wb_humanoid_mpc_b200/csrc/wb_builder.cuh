// Device-side instance builder: what SolverBase::preRun and the top of SqpSolver::runImpl produce on the host for every
// MPC instance -- mode schedule from the gait template, swing-z / impact-proximity references, arm-swing phase, target trajectories from the
// velocity command, the event-aware time grid, the initial guess (WeightCompInitializer, or the previous solution shifted) -- computed on the
// GPU from a few numbers per instance (x0, gait id, gait start time, velocity command), so that a batch of MPC cycles needs ~0.5 kB of host
// data per instance instead of the 150 kB of per-node arrays of b200sqp_upload_instances.
//
// Restates (function by function, same arithmetic order so that the results agree with the host layer to round-off):
//   GaitSchedule::insertModeSequenceTemplate / getModeSchedule / tileModeSequenceTemplate   humanoid_common_mpc/src/gait/GaitSchedule.cpp:60-140
//   SwingTrajectoryPlanner::update, getZ*Constraint, getImpactProximityFactor                humanoid_common_mpc/src/swing_foot_planner/SwingTrajectoryPlanner.cpp
//   SplineCpg / CubicSpline                                                                  .../swing_foot_planner/SplineCpg.cpp, CubicSpline.cpp:38-85
//   SwitchedModelReferenceManager::getContactFlags / getPhaseVariable                        .../reference_manager/SwitchedModelReferenceManager.cpp:54-135
//   WBMpcTargetTrajectoriesCalculator::commandedVelocityToTargetTrajectories                 humanoid_wb_mpc/src/command/WBMpcTargetTrajectoriesCalculator.cpp:82-136
//   timeDiscretizationWithEvents                                                             ocs2_oc/src/oc_data/TimeDiscretization.cpp:60-114
//   initializeStateInputTrajectories + WeightCompInitializer + LinearInterpolation           ocs2_oc/src/multiple_shooting/Initialization.cpp:35-79
// through this repository's host restatement wb_humanoid_mpc_b200/host/references.hpp, against which tests/test_gpu_builder.py compares every array.
// The warm start interpolates the previous primal solution as it is (toPrimalSolution time stamps); the reference's trajectorySpread
// re-stamping (SqpSolver.cpp:211-213) is a host-layer feature and is not applied here.
#pragma once
#include "wb_solver.cuh"

namespace b200sqp {

constexpr int BLD_MAX_EVENTS = 192;   // events of one mode-schedule window [t0 - T, tf + T]
constexpr int BLD_MAX_GAIT_MODES = B200SQP_MAX_GAIT_MODES;
constexpr double kLimitEpsB = 1e-6;   // numeric_traits::limitEpsilon
enum BldMode { M_FLY = 0, M_RF = 1, M_LF = 2, M_STANCE = 3 };   // humanoid_common_mpc/gait/MotionPhaseDefinition.h

struct BldSched {   // one instance's mode schedule window: modes[j] is active before events[j]; modes[n] after the last event
  int n;
  double ev[BLD_MAX_EVENTS];
  signed char mode[BLD_MAX_EVENTS + 1];
};
struct BldScratch {   // per instance, in global memory
  BldSched sched;
  double ttTime[3], ttState[3][NX];   // TargetTrajectories (3 knots)
  int nNodes, overflow;
};
struct BldDev {   // device pointers of the builder (allocated with the batch)
  BldScratch* scratch;   // [B]
  double* gridT;         // [B][cap]
  uint8_t* gridE;        // [B][cap]
  const double* cmd;     // [B][4]
  const int* gait;       // [B]
  const double* gaitStart;   // [B]
  double *prevX, *prevU, *prevT;   // previous solution (warm start): strides prevN + 1 / prevN
  uint8_t* prevE;
  int cap, prevN, warm;
  double t0, tf, dt;
  b200sqp_builder_desc desc;
};

HD bool bldContact(int mode, int leg) { return leg == 0 ? (mode == M_LF || mode == M_STANCE) : (mode == M_RF || mode == M_STANCE); }   // modeNumber2StanceLeg
HD int bldLowerBound(const double* v, int n, double t) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (v[mid] < t) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}
HD int bldUpperBound(const double* v, int n, double t) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (!(t < v[mid])) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}
HD int bldModeAt(const BldSched& s, double t) { return s.mode[bldLowerBound(s.ev, s.n, t)]; }

// GaitSchedule (functional form of the host's stateful object): events e_0 = S, e_{j+1} = e_j + duration[j mod m] accumulated from S exactly as
// tileModeSequenceTemplate does, whole cycles until the last event >= ub; STANCE before e_0 and after the last event; of the events before
// lb all but the last are dropped and the mode before the first kept event is STANCE (getModeSchedule's trimming).
HD bool bldModeSchedule(const b200sqp_builder_desc& d, int gait, double start, double lb, double ub, BldSched& s) {
  const int m = d.gait_n_modes[gait];
  double e = start;
  int j = 0;        // index of event e in the untrimmed list
  int kept = 0;     // number of events written
  int modeBefore = M_STANCE;   // mode before event j
  bool haveOneBefore = false;
  double lastBefore = 0.0;
  int lastBeforeMode = M_STANCE;
  s.n = 0;
  // events < lb: remember only the last one
  for (;;) {
    const bool cycleEnd = (j % m == 0);   // `while (ev.back() < finalTime)` is tested before every whole cycle, the first included
    if (cycleEnd && !(e < ub)) {
      // tiling stops after a whole cycle once its last event reaches ub
      if (e < lb) {   // (degenerate: nothing inside the window)
        haveOneBefore = true;
        lastBefore = e;
        lastBeforeMode = modeBefore;
      } else {
        if (haveOneBefore && kept == 0) {
          s.ev[kept] = lastBefore;
          s.mode[kept] = M_STANCE;
          ++kept;
          haveOneBefore = false;
        }
        if (kept >= BLD_MAX_EVENTS) return false;
        s.ev[kept] = e;
        s.mode[kept] = (kept == 0) ? M_STANCE : static_cast<signed char>(modeBefore);
        ++kept;
      }
      break;
    }
    if (e < lb) {
      haveOneBefore = true;
      lastBefore = e;
      lastBeforeMode = modeBefore;
    } else {
      if (haveOneBefore && kept == 0) {
        s.ev[kept] = lastBefore;
        s.mode[kept] = M_STANCE;   // seq[0] = STANCE after trimming
        ++kept;
        haveOneBefore = false;
      }
      if (kept >= BLD_MAX_EVENTS) return false;
      s.ev[kept] = e;
      s.mode[kept] = (kept == 0) ? M_STANCE : static_cast<signed char>(modeBefore);
      ++kept;
    }
    modeBefore = d.gait_modes[gait][j % m];
    e = e + (d.gait_switching_times[gait][j % m + 1] - d.gait_switching_times[gait][j % m]);
    ++j;
  }
  (void)lastBeforeMode;
  if (kept == 0) {   // every event lies before lb: the host keeps the last one
    s.ev[0] = lastBefore;
    s.mode[0] = M_STANCE;
    kept = 1;
  }
  s.n = kept;
  s.mode[kept] = M_STANCE;
  return true;
}

// CubicSpline (CubicSpline.cpp:38-85) evaluated at t: out = {position, velocity, acceleration}
HD void bldCubic(double ts, double ps, double vs, double te, double pe, double ve, double t, double out[3]) {
  const double dt = te - ts;
  const double dp = pe - ps, dv = ve - vs;
  const double c0 = ps, c1 = vs * dt, c2 = -(3.0 * vs + dv) * dt + 3.0 * dp, c3 = (2.0 * vs + dv) * dt - 2.0 * dp;
  const double tn = (t - ts) / dt;
  out[0] = c3 * tn * tn * tn + c2 * tn * tn + c1 * tn + c0;
  out[1] = (3.0 * c3 * tn * tn + 2.0 * c2 * tn + c1) / dt;
  out[2] = (6.0 * c3 * tn + 2.0 * c2) / (dt * dt);
}
// SplineCpg(liftOff, midHeight, touchDown)
HD void bldCpg(double ts, double ps, double vs, double mid, double te, double pe, double ve, double t, double out[3]) {
  const double tm = (ts + te) / 2;
  if (t < tm) bldCubic(ts, ps, vs, tm, mid, 0.0, t, out);
  else bldCubic(tm, mid, 0.0, te, pe, ve, t, out);
}
// SwingTrajectoryPlanner::update + zReference + impactProximityFactor for one leg at time t: z = {pos, vel, acc}, ipf
HD void bldSwing(const b200sqp_builder_desc& d, const BldSched& s, int leg, double t, double z[3], double& ipf) {
  const double liftOffVelocity = d.swing[0], touchDownVelocity = d.swing[1], swingHeight = d.swing[2], touchDownHeightOffset = d.swing[3],
               swingTimeScale = d.swing[4], ipfLiftOffVelocity = d.swing[5], ipfTouchDownVelocity = d.swing[6], midV = d.swing[7];
  const int nph = s.n + 1;
  const int p = bldLowerBound(s.ev, s.n, t);
  const double liftH = 0.0, touchH = 0.0 + touchDownHeightOffset;
  if (bldContact(s.mode[p], leg)) {
    z[0] = liftH;   // SplineCpg({0, liftH, 0}, liftH, {1, liftH, 0}): constant
    z[1] = 0.0;
    z[2] = 0.0;
    ipf = 1.0;
    // (the host evaluates the constant spline at t, which returns exactly liftH / 1 only up to round-off: reproduce the evaluation)
    double o[3];
    bldCpg(0.0, liftH, 0.0, liftH, 1.0, liftH, 0.0, t, o);
    z[0] = o[0];
    z[1] = o[1];
    z[2] = o[2];
    bldCpg(0.0, 1.0, 0.0, 1.0, 1.0, 1.0, 0.0, t, o);
    ipf = o[0];
    return;
  }
  int start = -1, fin = nph - 1;
  for (int ip = p - 1; ip >= 0; --ip)
    if (bldContact(s.mode[ip], leg)) {
      start = ip;
      break;
    }
  for (int ip = p + 1; ip < nph; ++ip)
    if (bldContact(s.mode[ip], leg)) {
      fin = ip - 1;
      break;
    }
  // (the schedule starts and ends in STANCE, so both searches succeed)
  const double ts = s.ev[start < 0 ? 0 : start], te = s.ev[fin >= s.n ? s.n - 1 : fin];
  const bool prevC = bldContact(s.mode[p - 1], leg), nextC = bldContact(s.mode[p + 1], leg);
  double o[3];
  if (prevC && nextC) {
    const double sc = fmin(1.0, (te - ts) / swingTimeScale);
    bldCpg(ts, liftH, sc * liftOffVelocity, fmin(liftH, touchH) + sc * swingHeight, te, touchH, sc * touchDownVelocity, t, z);
    bldCpg(ts, 1.0, sc * ipfLiftOffVelocity, midV, te, 1.0, sc * ipfTouchDownVelocity, t, o);
  } else if (prevC) {
    const double mid = liftH + swingHeight;
    bldCpg(ts, liftH, liftOffVelocity, mid, te, mid, 0.0, t, z);
    bldCpg(ts, 1.0, ipfLiftOffVelocity, midV, te, midV, 0.0, t, o);
  } else if (nextC) {
    const double mid = touchH + swingHeight;
    bldCpg(ts, mid, 0.0, mid, te, touchH, touchDownVelocity, t, z);
    bldCpg(ts, midV, 0.0, midV, te, 1.0, ipfTouchDownVelocity, t, o);
  } else {
    const double mid = touchH + swingHeight;
    bldCpg(ts, mid, 0.0, mid, te, mid, 0.0, t, z);
    bldCpg(ts, midV, 0.0, midV, te, midV, 0.0, t, o);
  }
  ipf = o[0];
}
// SwitchedModelReferenceManager::getPhaseVariable
HD double bldPhaseVariable(const BldSched& s, double t) {
  const int it = bldUpperBound(s.ev, s.n, t);
  const double nxt = it < s.n ? s.ev[it] : s.ev[s.n - 1];
  const double prv = it > 0 ? s.ev[it - 1] : s.ev[0] - 1.0;
  const int m = bldModeAt(s, t);
  if (m == M_LF) return 0.5 * (t - prv) / (nxt - prv);
  if (m == M_RF) return 0.5 + 0.5 * (t - prv) / (nxt - prv);
  return bldModeAt(s, prv - 0.01) == M_LF ? 0.5 : 0.0;
}
// LinearInterpolation::interpolate index / weight (zero-order extrapolation, lower range for duplicated times, tiny intervals snap)
HD void bldInterpIndex(const double* times, int n, double t, int& idx, double& alpha) {
  if (n <= 1) {
    idx = 0;
    alpha = 1.0;
    return;
  }
  const int index = bldLowerBound(times, n, t) - 1;
  const int last = n - 1;
  if (index < 0) {
    idx = 0;
    alpha = 1.0;
  } else if (index < last) {
    const double length = times[index + 1] - times[index], tillNext = times[index + 1] - t;
    idx = index;
    alpha = (length > 2.0 * kWeakEps) ? tillNext / length : (tillNext < 0.5 * length ? 0.0 : 1.0);
  } else {
    idx = last - 1 > 0 ? last - 1 : 0;
    alpha = 0.0;
  }
}

// ---- step 1, one work item per instance: schedule, target knots, time grid ---------------------------------------------------------------------------
HD void bldGridInstance(const WbDev& d, const BldDev& b, int i) {
  BldScratch& sc = b.scratch[i];
  const double T = b.tf - b.t0;
  sc.overflow = 0;
  const int gait = b.gait[i];
  if (!bldModeSchedule(b.desc, gait, b.gaitStart[i], b.t0 - T, b.tf + T, sc.sched)) sc.overflow = 1;
  // commandedVelocityToTargetTrajectories
  {
    const double* x0 = d.x0 + static_cast<size_t>(i) * NX;
    const double* cmd = b.cmd + 4 * i;
    double pose[6] = {x0[0], x0[1], x0[2], x0[3], 0.0, 0.0};
    const double yaw = pose[3];
    const double vgx = cos(yaw) * cmd[0] - sin(yaw) * cmd[1], vgy = sin(yaw) * cmd[0] + cos(yaw) * cmd[1];
    const double baseVel[6] = {vgx, vgy, 0.0, cmd[3], 0.0, 0.0};
    const double tMid = 0.7 * T;
    const double avg[3] = {(x0[NV] + vgx) / 2, (x0[NV + 1] + vgy) / 2, (x0[NV + 5] + cmd[3]) / 2};
    pose[2] = cmd[2];
    double mid[6], fin[6];
    mid[0] = pose[0] + avg[0] * tMid;
    mid[1] = pose[1] + avg[1] * tMid;
    mid[2] = cmd[2];
    mid[3] = pose[3] + avg[2] * tMid;
    mid[4] = mid[5] = 0.0;
    fin[0] = mid[0] + vgx * (T - tMid);
    fin[1] = mid[1] + vgy * (T - tMid);
    fin[2] = cmd[2];
    fin[3] = mid[3] + cmd[3] * (T - tMid);
    fin[4] = fin[5] = 0.0;
    sc.ttTime[0] = b.t0;
    sc.ttTime[1] = b.t0 + tMid;
    sc.ttTime[2] = b.t0 + T;
    const double* knots[3] = {pose, mid, fin};
    for (int k = 0; k < 3; ++k) {
      for (int j = 0; j < NX; ++j) sc.ttState[k][j] = 0.0;
      for (int j = 0; j < 6; ++j) {
        sc.ttState[k][j] = knots[k][j];
        sc.ttState[k][NV + j] = baseVel[j];
      }
      for (int j = 0; j < NJ; ++j) sc.ttState[k][6 + j] = b.desc.default_joint_state[j];
    }
  }
  // timeDiscretizationWithEvents (TimeDiscretization.cpp:60-114), then the PreEvent nodes are duplicated as PostEvent nodes
  {
    double* tOut = b.gridT + static_cast<size_t>(i) * b.cap;
    uint8_t* eOut = b.gridE + static_cast<size_t>(i) * b.cap;
    const BldSched& s = sc.sched;
    const double dtMin = 10.0 * kLimitEpsB;
    // pass 1 in place (without the duplicates): tOut / eOut hold `td`
    int n = 1;
    tOut[0] = b.t0;
    eOut[0] = 0;
    int nextIdx = bldLowerBound(s.ev, s.n, b.t0);
    double nextT = b.t0;
    int nextE = 0;
    while (tOut[n - 1] < b.tf) {
      nextT = nextT + b.dt;
      nextE = 0;
      if (nextIdx < s.n && nextT >= s.ev[nextIdx]) {
        nextT = s.ev[nextIdx];
        nextE = 1;
        ++nextIdx;
      }
      if (nextT >= b.tf) {
        nextT = b.tf;
        nextE = 0;
      }
      if (nextT > tOut[n - 1] + dtMin) {
        if (n >= b.cap) {
          sc.overflow = 1;
          break;
        }
        tOut[n] = nextT;
        eOut[n] = static_cast<uint8_t>(nextE);
        ++n;
      } else {
        tOut[n - 1] = nextT;
        eOut[n - 1] = static_cast<uint8_t>(nextE);
      }
    }
    if (eOut[0] == 1) eOut[0] = 2;
    // pass 2: duplicate the PreEvent nodes (from the back, in place)
    int nPre = 0;
    for (int k = 0; k < n; ++k) nPre += (eOut[k] == 1);
    const int total = n + nPre;
    if (total > b.cap) {
      sc.overflow = 1;
      sc.nNodes = total;
      return;
    }
    int w = total - 1;
    for (int k = n - 1; k >= 0; --k) {
      if (eOut[k] == 1) {
        tOut[w] = tOut[k];
        eOut[w] = 2;
        --w;
      }
      tOut[w] = tOut[k];
      eOut[w] = eOut[k];
      --w;
    }
    sc.nNodes = total;
  }
}

// ---- step 2, one group of nt work items per (node, instance): node data + initial guess ---------------------------------------------------------------
HD void bldNode(const WbDev& d, const BldDev& b, int k, int i, int tid, int nt) {
  const int N = d.N;
  if (k > N) return;
  const BldScratch& sc = b.scratch[i];
  const BldSched& s = sc.sched;
  const double* gt = b.gridT + static_cast<size_t>(i) * b.cap;
  const uint8_t* ge = b.gridE + static_cast<size_t>(i) * b.cap;
  const size_t node = static_cast<size_t>(i) * (N + 1) + k;
  const double tk = gt[k];
  const int ek = ge[k];
  const double t = tk + (ek == 2 ? kWeakEps : 0.0);   // getIntervalStart
  const int mode = bldModeAt(s, t);
  if (tid == 0) {
    d.t[node] = tk;
    d.event[node] = static_cast<uint8_t>(ek);
    d.arm[node] = sin(2.0 * 3.14159265358979323846 * (bldPhaseVariable(s, t) - 0.15));
  }
  if (tid < 2) {
    double z[3], ipf;
    bldSwing(b.desc, s, tid, t, z, ipf);
    d.contact[2 * node + tid] = bldContact(mode, tid) ? 1 : 0;
    for (int j = 0; j < 3; ++j) d.swing[(2 * node + tid) * 3 + j] = z[j];
    d.impact[2 * node + tid] = ipf;
  }
  // TargetTrajectories::getDesiredState: linear interpolation with clamping
  {
    int seg = -1;
    double a = 0.0;
    if (t <= sc.ttTime[0]) seg = -1;
    else if (t >= sc.ttTime[2]) seg = 3;
    else {
      seg = bldUpperBound(sc.ttTime, 3, t) - 1;
      a = (sc.ttTime[seg + 1] - t) / (sc.ttTime[seg + 1] - sc.ttTime[seg]);
    }
    for (int j = tid; j < NX; j += nt) {
      double v;
      if (seg < 0) v = sc.ttState[0][j];
      else if (seg >= 2) v = sc.ttState[2][j];
      else v = a * sc.ttState[seg][j] + (1 - a) * sc.ttState[seg + 1][j];
      d.xref[node * NX + j] = v;
    }
  }
  // ---- initializeStateInputTrajectories -----------------------------------------------------------------------------------------------------------
  const double* x0 = d.x0 + static_cast<size_t>(i) * NX;
  const bool warm = b.warm && b.prevN >= 1;
  const int pn = b.prevN + 1;   // samples of the previous primal solution
  const double* pT = b.prevT + static_cast<size_t>(i) * pn;
  const double* pX = b.prevX + static_cast<size_t>(i) * pn * NX;
  const double* pU = b.prevU + static_cast<size_t>(i) * b.prevN * NU;
  const uint8_t* pE = b.prevE + static_cast<size_t>(i) * pn;
  const double tStateTill = warm ? pT[pn - 1] : gt[0];
  const double tInputTill = warm ? pT[pn - 2] : gt[0];
  // state of node k: the state interpolated at the end of the last interval j < k that is covered by the previous solution (none: the state at the
  // initial time, or x0)
  {
    int src = -1;   // interval index whose end gives the state
    for (int j = k - 1; j >= 0; --j) {
      if (ge[j] == 1) continue;   // PreEvent node: x is carried over
      const double tj = gt[j] + (ge[j] == 2 ? kWeakEps : 0.0);
      const double tn = gt[j + 1] - (ge[j + 1] == 1 ? kWeakEps : 0.0);
      if (!(tj > tInputTill || tn > tStateTill)) {
        src = j;
        break;
      }
    }
    double tq;
    bool fromPrev;
    if (src >= 0) {
      tq = gt[src + 1] - (ge[src + 1] == 1 ? kWeakEps : 0.0);
      fromPrev = true;
    } else {
      tq = gt[0] + (ge[0] == 2 ? kWeakEps : 0.0);
      fromPrev = warm && tq < tStateTill;
    }
    if (fromPrev) {
      int idx;
      double alpha;
      bldInterpIndex(pT, pn, tq, idx, alpha);
      for (int j = tid; j < NX; j += nt) d.x[node * NX + j] = alpha * pX[static_cast<size_t>(idx) * NX + j] + (1.0 - alpha) * pX[static_cast<size_t>(idx + 1) * NX + j];
    } else {
      for (int j = tid; j < NX; j += nt) d.x[node * NX + j] = x0[j];
    }
  }
  if (k < N) {
    const size_t stage = static_cast<size_t>(i) * N + k;
    double* u = d.u + stage * NU;
    if (ek == 1) {
      for (int j = tid; j < NU; j += nt) u[j] = 0.0;
    } else {
      const double tn = gt[k + 1] - (ge[k + 1] == 1 ? kWeakEps : 0.0);
      if (t > tInputTill || tn > tStateTill) {
        // WeightCompInitializer: weight-compensating normal forces on the stance feet
        const bool l = bldContact(mode, 0), r = bldContact(mode, 1);
        const int ns = int(l) + int(r);
        const double fz = ns > 0 ? b.desc.total_mass * 9.81 / ns : 0.0;
        for (int j = tid; j < NU; j += nt) u[j] = (j == 2 && l) ? fz : ((j == 8 && r) ? fz : 0.0);
      } else {
        // inputs of the previous primal solution (toPrimalSolution): sample m < prevN is u_m, except that a PreEvent node repeats the input
        // before it; the last sample repeats the last input
        int idx;
        double alpha;
        bldInterpIndex(pT, pn, t, idx, alpha);
        auto inputOf = [&](int m2) {
          if (m2 >= b.prevN) m2 = b.prevN - 1;
          while (m2 > 0 && pE[m2] == 1) --m2;
          return pU + static_cast<size_t>(m2) * NU;
        };
        const double* ua = inputOf(idx);
        const double* ub = inputOf(idx + 1);
        for (int j = tid; j < NU; j += nt) u[j] = alpha * ua[j] + (1.0 - alpha) * ub[j];
      }
    }
  }
}

// Perfect-tracking closed loop (the dummy simulation of the reference's launch files, WBMpcRobotSim.cpp: the next measured state is the planned
// one): x0 = the previous primal solution interpolated at the new initial time, on the device.
HD void bldX0FromPrevious(const WbDev& d, const BldDev& b, int i, int tid, int nt) {
  const int pn = b.prevN + 1;
  const double* pT = b.prevT + static_cast<size_t>(i) * pn;
  const double* pX = b.prevX + static_cast<size_t>(i) * pn * NX;
  int idx;
  double alpha;
  bldInterpIndex(pT, pn, b.t0, idx, alpha);
  for (int j = tid; j < NX; j += nt) d.x0[static_cast<size_t>(i) * NX + j] = alpha * pX[static_cast<size_t>(idx) * NX + j] + (1.0 - alpha) * pX[static_cast<size_t>(idx + 1) * NX + j];
}

#ifdef __CUDACC__
__global__ void builder_x0_kernel(WbDev d, BldDev b) { bldX0FromPrevious(d, b, blockIdx.x, threadIdx.x, blockDim.x); }
__global__ void builder_grid_kernel(WbDev d, BldDev b) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < d.B) bldGridInstance(d, b, i);
}
__global__ void __launch_bounds__(64) builder_nodes_kernel(WbDev d, BldDev b) {
  bldNode(d, b, blockIdx.x, blockIdx.y, threadIdx.x, blockDim.x);
}
#endif

}  // namespace b200sqp
