"""Host-side instance builder: what SolverBase::preRun + runImpl's bookkeeping hand to the LQ stage, as flat per-node arrays.

Restates (paths relative to /root/reference/):
  GaitSchedule                 humanoid_nmpc/humanoid_common_mpc/src/gait/GaitSchedule.cpp:46-139
  ModeSchedule::modeAtTime     lib/ocs2_ros2/ocs2_core/src/reference/ModeSchedule.cpp:48-51
  SwitchedModelReferenceManager::modifyReferences / getPhaseVariable / getContactFlags
                               humanoid_nmpc/humanoid_common_mpc/src/reference_manager/SwitchedModelReferenceManager.cpp:54-154
  SwingTrajectoryPlanner / SplineCpg / CubicSpline
                               humanoid_nmpc/humanoid_common_mpc/src/swing_foot_planner/SwingTrajectoryPlanner.cpp:50-271,
                               SplineCpg.cpp:38-63, CubicSpline.cpp:38-85
  timeDiscretizationWithEvents lib/ocs2_ros2/ocs2_oc/src/oc_data/TimeDiscretization.cpp:40-114
  commandedVelocityToTargetTrajectories
                               humanoid_nmpc/humanoid_wb_mpc/src/command/WBMpcTargetTrajectoriesCalculator.cpp:82-136
  WeightCompInitializer / initializeStateInputTrajectories (cold start)
                               humanoid_nmpc/humanoid_common_mpc/src/initialization/WeightCompInitializer.cpp:66-70,
                               lib/ocs2_ros2/ocs2_oc/src/multiple_shooting/Initialization.cpp:35-79
These run on the host for every MPC instance (SURVEY.md §8(f)-1 lists moving them to the device as the next step).
"""
from __future__ import annotations

import bisect
import math
from dataclasses import dataclass, field

import numpy as np

FLY, RF, LF, STANCE = 0, 1, 2, 3
MODE_NAMES = {"FLY": FLY, "RF": RF, "LF": LF, "STANCE": STANCE}
LIMIT_EPS, WEAK_EPS = 1e-6, 1e-9
EV_NONE, EV_PRE, EV_POST = 0, 1, 2


def mode_to_contacts(mode: int):
    """modeNumber2StanceLeg: {left, right}"""
    return [(False, False), (False, True), (True, False), (True, True)][mode]


@dataclass
class ModeSchedule:
    event_times: list = field(default_factory=list)
    mode_sequence: list = field(default_factory=lambda: [STANCE])

    def mode_at(self, t: float) -> int:
        return self.mode_sequence[bisect.bisect_left(self.event_times, t)]


class GaitSchedule:
    def __init__(self, init_events=(0.5,), init_modes=(STANCE, STANCE), template_modes=(STANCE,), template_times=(0.0, 0.5),
                 phase_transition_stance_time=0.0):
        self.ms = ModeSchedule(list(init_events), list(init_modes))
        self.t_modes, self.t_times = list(template_modes), list(template_times)
        self.ptst = phase_transition_stance_time

    def insert_template(self, modes, times, start_time, final_time):
        self.t_modes, self.t_times = list(modes), list(times)
        ev, seq = self.ms.event_times, self.ms.mode_sequence
        idx = bisect.bisect_left(ev, start_time)
        if idx < len(ev):
            del ev[idx:]
            del seq[idx + 1:]
        ptst = 0.0 if (seq and seq[-1] == STANCE) else self.ptst
        if ptst > 0.0:
            ev.append(start_time)
            seq.append(STANCE)
        self._tile(start_time + ptst, final_time)

    def get_mode_schedule(self, lower, upper) -> ModeSchedule:
        ev, seq = self.ms.event_times, self.ms.mode_sequence
        idx = bisect.bisect_left(ev, lower)
        if idx > 0:
            del ev[: idx - 1]
            del seq[: idx - 1]
            seq[0] = STANCE
        tiling_start = upper if not ev else ev[-1]
        del ev[-1:]
        del seq[-1:]
        self._tile(tiling_start, upper)
        return ModeSchedule(list(ev), list(seq))

    def _tile(self, start, final):
        ev, seq = self.ms.event_times, self.ms.mode_sequence
        if not self.t_modes:
            return
        if ev and start <= ev[-1]:
            raise RuntimeError("The initial time for template-tiling is not greater than the last event time.")
        ev.append(start)
        while ev[-1] < final:
            for i, m in enumerate(self.t_modes):
                seq.append(m)
                ev.append(ev[-1] + (self.t_times[i + 1] - self.t_times[i]))
        seq.append(STANCE)


class CubicSpline:
    def __init__(self, t0, p0, v0, t1, p1, v1):
        self.t0, self.dt = t0, t1 - t0
        dp, dv = p1 - p0, v1 - v0
        self.c0 = p0
        self.c1 = v0 * self.dt
        self.c2 = -(3.0 * v0 + dv) * self.dt + 3.0 * dp
        self.c3 = (2.0 * v0 + dv) * self.dt - 2.0 * dp

    def pos(self, t):
        tn = (t - self.t0) / self.dt
        return self.c3 * tn ** 3 + self.c2 * tn ** 2 + self.c1 * tn + self.c0

    def vel(self, t):
        tn = (t - self.t0) / self.dt
        return (3.0 * self.c3 * tn * tn + 2.0 * self.c2 * tn + self.c1) / self.dt

    def acc(self, t):
        tn = (t - self.t0) / self.dt
        return (6.0 * self.c3 * tn + 2.0 * self.c2) / (self.dt * self.dt)


class SplineCpg:
    def __init__(self, lift, mid_height, touch):
        (t0, p0, v0), (t1, p1, v1) = lift, touch
        self.mid = (t0 + t1) / 2
        self.left = CubicSpline(t0, p0, v0, self.mid, mid_height, 0.0)
        self.right = CubicSpline(self.mid, mid_height, 0.0, t1, p1, v1)

    def pos(self, t):
        return self.left.pos(t) if t < self.mid else self.right.pos(t)

    def vel(self, t):
        return self.left.vel(t) if t < self.mid else self.right.vel(t)

    def acc(self, t):
        return self.left.acc(t) if t < self.mid else self.right.acc(t)


class SwingTrajectoryPlanner:
    def __init__(self, cfg: dict, n_feet=2):
        self.c, self.n = cfg, n_feet

    def update(self, ms: ModeSchedule, terrain_height=0.0):
        c = self.c
        modes, ev = ms.mode_sequence, ms.event_times
        nph = len(modes)
        lift_h, touch_h = terrain_height, terrain_height + c["touchDownHeightOffset"]
        self.events = list(ev)
        self.height, self.impact = [], []
        for j in range(self.n):
            flags = [mode_to_contacts(m)[j] for m in modes]
            hs, ips = [], []
            for p in range(nph):
                if flags[p]:
                    hs.append(SplineCpg((0.0, lift_h, 0.0), lift_h, (1.0, lift_h, 0.0)))
                    ips.append(SplineCpg((0.0, 1.0, 0.0), 1.0, (1.0, 1.0, 0.0)))
                    continue
                start = next((ip for ip in range(p - 1, -1, -1) if flags[ip]), -1)
                final = next((ip - 1 for ip in range(p + 1, nph) if flags[ip]), nph - 1)
                if start < 0:
                    raise RuntimeError(f"The time of take-off for the first swing of the EE with ID {j} is not defined.")
                if final >= nph - 1:
                    raise RuntimeError(f"The time of touch-down for the last swing of the EE with ID {j} is not defined.")
                ts, tf = ev[start], ev[final]
                prev_c, next_c = flags[p - 1], flags[p + 1]
                mid_v = c["ipfMidPointValue"]
                if prev_c and next_c:
                    s = min(1.0, (tf - ts) / c["swingTimeScale"])
                    hs.append(SplineCpg((ts, lift_h, s * c["liftOffVelocity"]), min(lift_h, touch_h) + s * c["swingHeight"],
                                        (tf, touch_h, s * c["touchDownVelocity"])))
                    ips.append(SplineCpg((ts, 1.0, s * c["ipfLiftOffVelocity"]), mid_v, (tf, 1.0, s * c["ipfTouchDownVelocity"])))
                elif prev_c:
                    mid = lift_h + c["swingHeight"]
                    hs.append(SplineCpg((ts, lift_h, c["liftOffVelocity"]), mid, (tf, mid, 0.0)))
                    ips.append(SplineCpg((ts, 1.0, c["ipfLiftOffVelocity"]), mid_v, (tf, mid_v, 0.0)))
                elif next_c:
                    mid = touch_h + c["swingHeight"]
                    hs.append(SplineCpg((ts, mid, 0.0), mid, (tf, touch_h, c["touchDownVelocity"])))
                    ips.append(SplineCpg((ts, mid_v, 0.0), mid_v, (tf, 1.0, c["ipfTouchDownVelocity"])))
                else:
                    mid = touch_h + c["swingHeight"]
                    hs.append(SplineCpg((ts, mid, 0.0), mid, (tf, mid, 0.0)))
                    ips.append(SplineCpg((ts, mid_v, 0.0), mid_v, (tf, mid_v, 0.0)))
            self.height.append(hs)
            self.impact.append(ips)

    def _idx(self, t):
        return bisect.bisect_left(self.events, t)

    def z_ref(self, leg, t):
        s = self.height[leg][self._idx(t)]
        return s.pos(t), s.vel(t), s.acc(t)

    def impact_factor(self, leg, t):
        return self.impact[leg][self._idx(t)].pos(t)


def phase_variable(ms: ModeSchedule, t: float) -> float:
    ev = ms.event_times
    it = bisect.bisect_right(ev, t)
    # the reference dereferences both neighbours unchecked (SwitchedModelReferenceManager.cpp:63-65); outside the event range the phase is a
    # STANCE phase and only `prv` is used, for a mode lookup: any time before the first event gives the same answer
    nxt = ev[it] if it < len(ev) else ev[-1]
    prv = ev[it - 1] if it > 0 else ev[0] - 1.0
    m = ms.mode_at(t)
    if m == LF:
        return 0.5 * (t - prv) / (nxt - prv)
    if m == RF:
        return 0.5 + 0.5 * (t - prv) / (nxt - prv)
    return 0.5 if ms.mode_at(prv - 0.01) == LF else 0.0


def time_discretization_with_events(t0, tf, dt, event_times, dt_min=10.0 * LIMIT_EPS):
    td = [[t0, EV_NONE]]
    nxt_idx = bisect.bisect_left(event_times, t0)
    nxt = [t0, EV_NONE]
    while td[-1][0] < tf:
        nxt = [nxt[0] + dt, EV_NONE]
        if nxt_idx < len(event_times) and nxt[0] >= event_times[nxt_idx]:
            nxt = [event_times[nxt_idx], EV_PRE]
            nxt_idx += 1
        if nxt[0] >= tf:
            nxt = [tf, EV_NONE]
        if nxt[0] > td[-1][0] + dt_min:
            td.append(list(nxt))
        else:
            td[-1] = list(nxt)
    if td[0][1] == EV_PRE:
        td[0][1] = EV_POST
    out = []
    for t, e in td:
        out.append((t, e))
        if e == EV_PRE:
            out.append((t, EV_POST))
    return np.array([t for t, _ in out]), np.array([e for _, e in out], dtype=np.uint8)


def interval_start(t, e):
    return t + (WEAK_EPS if e == EV_POST else 0.0)


def velocity_command_targets(model: dict, t0, x0, cmd, horizon):
    """3-knot TargetTrajectories for cmd = [v_x, v_y, pelvis height, yaw rate] (steady-state command filter)."""
    nj = model["nj"]
    nv = 6 + nj
    pose = np.array(x0[:6], float)
    pose[4] = pose[5] = 0.0
    yaw = pose[3]
    vg = np.array(cmd, float)
    vg[0] = math.cos(yaw) * cmd[0] - math.sin(yaw) * cmd[1]
    vg[1] = math.sin(yaw) * cmd[0] + math.cos(yaw) * cmd[1]
    base_vel = np.array([vg[0], vg[1], 0.0, vg[3], 0.0, 0.0])
    t_mid = 0.7 * horizon
    bv = x0[nv:nv + 6]
    avg = np.array([(bv[0] + vg[0]) / 2, (bv[1] + vg[1]) / 2, (bv[5] + vg[3]) / 2])
    pose[2] = vg[2]

    def integrate(p, av, h, dT):
        q = p.copy()
        q[0] += av[0] * dT
        q[1] += av[1] * dT
        q[2] = h
        q[3] += av[2] * dT
        q[4] = q[5] = 0.0
        return q

    mid = integrate(pose, avg, vg[2], t_mid)
    fin = integrate(mid, np.array([vg[0], vg[1], vg[3]]), vg[2], horizon - t_mid)
    joints = np.array(model["reference"]["defaultJointState"])
    states = [np.concatenate([p, joints, base_vel, np.zeros(nj)]) for p in (pose, mid, fin)]
    return np.array([t0, t0 + t_mid, t0 + horizon]), np.array(states)


def velocity_command_targets_centroidal(model: dict, t0, x0, cmd, horizon, base_vel=None):
    """CentroidalMpcTargetTrajectoriesCalculator::commandedVelocityToTargetTrajectories (humanoid_centroidal_mpc/src/command/
    CentroidalMpcTargetTrajectoriesCalculator.cpp:82-160).  base_vel = Ab^-1 * x0[:6] (the reference multiplies the inverse base block of
    the centroidal momentum matrix with the NORMALIZED momentum, i.e. flow_map(x0, u = 0)[6:12] / mass); it may be omitted for x0[:6] = 0.
    The yaw average uses base_vel[5] -- the roll rate in the ZYX ordering -- as the reference does."""
    nj = model["nj"]
    x0 = np.asarray(x0, float)
    if base_vel is None:
        if np.any(x0[:6] != 0.0):
            raise ValueError("base_vel = Ab^-1 * x0[:6] is needed for a non-zero initial momentum (centroidal.base_velocity)")
        base_vel = np.zeros(6)
    pose = np.array(x0[6:12], float)
    pose[4] = pose[5] = 0.0
    yaw = pose[3]
    vg = np.array(cmd, float)
    vg[0] = math.cos(yaw) * cmd[0] - math.sin(yaw) * cmd[1]
    vg[1] = math.sin(yaw) * cmd[0] + math.cos(yaw) * cmd[1]
    momentum = np.array([vg[0], vg[1], 0.0, 0.0, 0.0, vg[3] / sum(model["mass"])])
    t_mid = 0.7 * horizon
    avg = np.array([(base_vel[0] + vg[0]) / 2, (base_vel[1] + vg[1]) / 2, (base_vel[5] + vg[3]) / 2])
    pose[2] = vg[2]

    def integrate(p, av, h, dT):
        q = p.copy()
        q[0] += av[0] * dT
        q[1] += av[1] * dT
        q[2] = h
        q[3] += av[2] * dT
        q[4] = q[5] = 0.0
        return q

    mid = integrate(pose, avg, vg[2], t_mid)
    fin = integrate(mid, np.array([vg[0], vg[1], vg[3]]), vg[2], horizon - t_mid)
    joints = np.array(model["reference"]["defaultJointState"])
    states = [np.concatenate([momentum, p, joints]) for p in (pose, mid, fin)]
    return np.array([t0, t0 + t_mid, t0 + horizon]), np.array(states)


def interp_targets(times, states, t):
    """LinearInterpolation::interpolate with clamping (TargetTrajectories::getDesiredState)"""
    if t <= times[0]:
        return states[0].copy()
    if t >= times[-1]:
        return states[-1].copy()
    i = bisect.bisect_right(list(times), t) - 1
    a = (times[i + 1] - t) / (times[i + 1] - times[i])
    return a * states[i] + (1 - a) * states[i + 1]


def weight_compensating_input(model: dict, contacts):
    u = np.zeros(model["nu"])
    ns = int(contacts[0]) + int(contacts[1])
    if ns > 0:
        fz = sum(model["mass"]) * 9.81 / ns
        for c in range(2):
            if contacts[c]:
                u[6 * c + 2] = fz
    return u


def linear_interpolate(t, times, data):
    """LinearInterpolation::interpolate (ocs2_core/include/ocs2_core/misc/implementation/LinearInterpolation.h:67-129): zero-order
    extrapolation, for duplicated times the lower range ( ] is selected, tiny intervals snap to the closest sample."""
    times = list(times)
    if len(times) <= 1:
        return np.array(data[0], float).copy()
    index = bisect.bisect_left(times, t) - 1          # lookup::findIntervalInTimeArray
    last = len(times) - 1
    if index < 0:
        idx, alpha = 0, 1.0
    elif index < last:
        length, till_next = times[index + 1] - times[index], times[index + 1] - t
        if length > 2.0 * WEAK_EPS:
            idx, alpha = index, till_next / length
        else:
            idx, alpha = index, (0.0 if till_next < 0.5 * length else 1.0)
    else:
        idx, alpha = max(last - 1, 0), 0.0
    return alpha * np.asarray(data[idx], float) + (1.0 - alpha) * np.asarray(data[idx + 1], float)


def to_primal_solution(t_nodes, events, x, u, mode_schedule=None):
    """multiple_shooting::toPrimalSolution (ocs2_oc/src/multiple_shooting/Helpers.cpp:60-82): inputs at PreEvent nodes repeat the previous
    input and the last input is repeated so that time, state and input trajectories have equal length."""
    u = [np.array(r, float) for r in u]
    for i in range(len(u)):
        if events[i] == EV_PRE and i > 0:
            u[i] = u[i - 1].copy()
    u.append(u[-1].copy())
    out = dict(t=np.array(t_nodes, float), x=np.array(x, float), u=np.array(u))
    if mode_schedule is not None:
        out["mode_schedule"] = mode_schedule
    return out


class TrajectorySpreading:
    """ocs2::TrajectorySpreading (ocs2_oc/src/trajectory_adjustment/TrajectorySpreading.cpp:49-166,268-369 and the templates in
    include/ocs2_oc/trajectory_adjustment/TrajectorySpreading.h:124-183): adapts trajectories computed for an old mode schedule to a new one
    by spreading the values next to the moved event times and truncating where the mode sequences stop matching."""

    def set(self, old: ModeSchedule, new: ModeSchedule, old_time):
        ub = lambda a, t: bisect.bisect_right(list(a), t)
        lb = lambda a, t: bisect.bisect_left(list(a), t)
        t0, tf = old_time[0], old_time[-1]
        old_first, old_last = ub(old.event_times, t0), ub(old.event_times, tf)
        new_first, new_last = ub(new.event_times, t0), ub(new.event_times, tf)
        old_start, new_start = old_first, new_first
        w = 0
        while old_start < len(old.mode_sequence):
            a = old.mode_sequence[old_start: old_last + 1]
            b = new.mode_sequence[new_start: new_last + 1]
            w = 0
            while w < len(a) and w < len(b) and a[w] == b[w]:      # std::mismatch
                w += 1
            if w > 0:
                break
            old_start += 1
        old_m, new_m = [], []
        self.keep_event_data = (0, 0)
        if w > 0:
            old_m = list(old.event_times[old_start: old_start + w - 1])
            new_m = list(new.event_times[new_start: new_start + w - 1])
            self.keep_event_data = (old_start - old_first, old_start - old_first + w - 1)
        if w > 0 and old_start > old_first:
            old_m.insert(0, old.event_times[old_start - 1])
            new_m.insert(0, t0 - 1e-4)
        old_last_matched = (old_start + w - 1 == old_last)
        new_last_matched = (new_start + w - 1 == new_last)
        if w > 0 and (not old_last_matched) and (new_last_matched or old.event_times[old_start + w - 1] < new.event_times[new_start + w - 1]):
            old_m.append(old.event_times[old_start + w - 1])
            new_m.append(tf + 1e-4 if new_last_matched else new.event_times[new_start + w - 1])
        self.erase_from = len(old_time)
        if w == 0:
            self.erase_from = 0
        elif not new_last_matched:
            self.erase_from = lb(old_time, new.event_times[new_start + w - 1])
        self._strategy(list(old_time), old_m, new_m)
        self.will_truncate = self.erase_from < len(old_time)
        self.will_spread = len(self.value_idx) > 0
        return self

    @staticmethod
    def _post_event_indices(event_times, time):
        out = []
        for i, e in enumerate(event_times):
            if i == len(event_times) - 1 and e == time[-1]:
                out.append(len(time) - 1)
            else:
                out.append(bisect.bisect_right(time, e))
        return out

    def _strategy(self, old_time, old_m, new_m):
        self.begin, self.end, self.value_idx, self.post_event_indices, self.matched_event_times = [], [], [], [], []
        old_post = self._post_event_indices(old_m, old_time)
        new_post = self._post_event_indices(new_m, old_time)
        for j in range(len(old_post)):
            if new_post[j] < old_post[j]:        # backward spreading
                self.begin.append(new_post[j])
                self.end.append(min(old_post[j], self.erase_from))
                self.value_idx.append(old_post[j])
            elif new_post[j] > old_post[j]:      # forward spreading
                self.begin.append(old_post[j] if j == 0 else max(old_post[j], new_post[j - 1]))
                self.end.append(new_post[j])
                self.value_idx.append(old_post[j] - 1)
            if new_post[j] != 0 and new_post[j] < self.erase_from:
                self.post_event_indices.append(new_post[j])
                self.matched_event_times.append(new_m[j])

    def adjust_trajectory(self, traj):
        traj = [np.array(v, float) for v in traj][: self.erase_from]
        values = [traj[i].copy() for i in self.value_idx]
        for b, e, v in zip(self.begin, self.end, values):
            for j in range(b, e):
                traj[j] = v.copy()
        return traj

    def adjust_time_trajectory(self, time):
        time = list(time)[: self.erase_from]
        for i, te in zip(self.post_event_indices, self.matched_event_times):
            time[i - 1] = te
            time[i] = min(te + WEAK_EPS, time[-1])
        return time


def trajectory_spread(old_ms: ModeSchedule, new_ms: ModeSchedule, primal: dict) -> dict:
    """trajectorySpread(oldModeSchedule, newModeSchedule, primalSolution)
    (include/ocs2_oc/trajectory_adjustment/TrajectorySpreadingHelperFunctions.h:124-143), called by SqpSolver::runImpl (SqpSolver.cpp:211-213)"""
    ts = TrajectorySpreading().set(old_ms, new_ms, primal["t"])
    out = dict(x=np.array(ts.adjust_trajectory(primal["x"])), u=np.array(ts.adjust_trajectory(primal["u"])),
               t=np.array(ts.adjust_time_trajectory(primal["t"])), post_event_indices=list(ts.post_event_indices), mode_schedule=new_ms,
               will_truncate=ts.will_truncate, will_spread=ts.will_spread)
    return out


def initialize_state_input_trajectories(model, x0, t_nodes, events, contact, previous=None):
    """multiple_shooting::initializeStateInputTrajectories (ocs2_oc/src/multiple_shooting/Initialization.cpp:35-79): interpolate the previous
    primal solution where it overlaps the new horizon, WeightCompInitializer for the tail (and for everything on a cold start)."""
    n = len(t_nodes)
    x0 = np.asarray(x0, float)
    t_state_till = t_input_till = t_nodes[0]
    if previous is not None and len(previous["t"]) >= 2:
        t_state_till, t_input_till = previous["t"][-1], previous["t"][-2]
    xs = []
    t_init = interval_start(t_nodes[0], events[0])
    xs.append(linear_interpolate(t_init, previous["t"], previous["x"]) if t_init < t_state_till else x0.copy())
    us = []
    for i in range(n - 1):
        if events[i] == EV_PRE:
            us.append(np.zeros(model["nu"]))
            xs.append(xs[-1].copy())
            continue
        t = interval_start(t_nodes[i], events[i])
        t_next = t_nodes[i + 1] - (WEAK_EPS if events[i + 1] == EV_PRE else 0.0)
        if t > t_input_till or t_next > t_state_till:
            us.append(weight_compensating_input(model, contact[i]))
            xs.append(xs[-1].copy())
        else:
            us.append(linear_interpolate(t, previous["t"], previous["u"]))
            xs.append(linear_interpolate(t_next, previous["t"], previous["x"]))
    return np.array(xs), np.array(us)


def build_instance(model: dict, x0, t0=0.0, horizon=None, dt=None, gait="stance", gait_start=None, cmd=None, previous=None, base_vel=None):
    """One MPC instance -> dict of per-node arrays in the layout of b200sqp_upload_instances.
    `previous` = to_primal_solution(...) of the last solve enables the reference's warm start; None is a cold start.
    model["kind"] == "centroidal" selects the centroidal target-trajectory rule (base_vel: see velocity_command_targets_centroidal); the
    schedule, swing planner, time grid and initializer are shared by both MPCs (CentroidalWeightCompInitializer keeps the momentum)."""
    sq = model["sqp"]
    horizon = sq["timeHorizon"] if horizon is None else horizon
    dt = sq["dt"] if dt is None else dt
    tf = t0 + horizon
    x0 = np.asarray(x0, float)
    gs = GaitSchedule()
    if gait != "stance":
        g = model["gaits"][gait]
        start = t0 if gait_start is None else gait_start
        gs.insert_template([MODE_NAMES[m] for m in g["modeSequence"]], g["switchingTimes"], start, tf + horizon)
    ms = gs.get_mode_schedule(t0 - horizon, tf + horizon)
    planner = SwingTrajectoryPlanner(model["swing"])
    planner.update(ms, 0.0)
    t_nodes, events = time_discretization_with_events(t0, tf, dt, ms.event_times)
    n = len(t_nodes)
    cmd = [0.0, 0.0, model["reference"]["defaultBaseHeight"], 0.0] if cmd is None else list(cmd)
    if model.get("kind") == "centroidal":
        tt, ts = velocity_command_targets_centroidal(model, t0, x0, cmd, horizon, base_vel)
    else:
        tt, ts = velocity_command_targets(model, t0, x0, cmd, horizon)
    nx, nu = model["nx"], model["nu"]
    contact = np.zeros((n, 2), dtype=np.uint8)
    swing = np.zeros((n, 2, 3))
    impact = np.ones((n, 2))
    arm = np.zeros(n)
    xref = np.zeros((n, nx))
    for i in range(n):
        t = interval_start(t_nodes[i], events[i])
        m = ms.mode_at(t)
        contact[i] = mode_to_contacts(m)
        for c in range(2):
            swing[i, c] = planner.z_ref(c, t)
            impact[i, c] = planner.impact_factor(c, t)
        arm[i] = math.sin(2 * math.pi * (phase_variable(ms, t) - 0.15))
        xref[i] = interp_targets(tt, ts, t)
    if previous is not None and previous.get("mode_schedule") is not None:
        previous = trajectory_spread(previous["mode_schedule"], ms, previous)   # SqpSolver.cpp:211-213
    x_init, u_init = initialize_state_input_trajectories(model, x0, t_nodes, events, contact, previous)
    return dict(x0=x0, x_init=x_init, u_init=u_init, t_nodes=t_nodes, node_event=events, contact_flags=contact, swing_ref=swing,
                impact_factor=impact, arm_phase=arm, x_ref=xref, mode_schedule=ms)
