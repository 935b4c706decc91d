"""Pins the oracle's TrajectorySpreading restatement (oracle/trajectory_spreading.hpp) on the reference's own recipes, then pins the
product's two host restatements (C++ host/references.hpp, Python references.py) on the oracle -- SURVEY.md section 8 row a3.

Reference test: lib/ocs2_ros2/ocs2_oc/test/trajectory_adjustment/TrajectorySpreadingTest.cpp -- all 17 TEST_F cases (:259-516), the property
checks of its checkResults() (:178-245) and the Status flags the cases EXPECT.  The reference rolls a linear system out with
TimeTriggeredRollout; only the time stamps and the mode of each sample enter the checks, so the roll-out is restated as its time grid:
integration samples every 0.01 s, a pre-event sample at every event time and a post-event sample eps later (RolloutBase.cpp: later intervals
start at eventTime + weakEpsilon)."""
import bisect

import numpy as np
import pytest

import oracle_lib as orc

EPS = 1e-9


def mode_at_time(ev, modes, t, final=False):
    """ModeSchedule::modeAtTime = lower_bound (an event time belongs to the pre-event mode); the test's modeAtTime uses upper_bound at the final time"""
    return modes[bisect.bisect_right(ev, t) if final else bisect.bisect_left(ev, t)]


def rollout(ev, modes, period):
    """time trajectory + mode per sample + post-event indices + per-event data (the pre-event mode, TrajectorySpreadingTest.cpp:107-123).
    Intervals as RolloutBase::findActiveModesTimeInterval (RolloutBase.cpp:44-68) builds them -- events in (t0, tf], an event at the final
    time included, EVERY interval (the first too) starting eps after its switching time -- and samples as TimeTriggeredRollout::runImpl
    (TimeTriggeredRollout.cpp:85-110) concatenates them: a degenerate interval contributes one sample at its end time."""
    t0, tf = period
    bounds = [t0] + [e for e in ev if t0 < e <= tf] + [tf]
    t, post = [], []
    for i in range(len(bounds) - 1):
        a, b = min(bounds[i] + EPS, bounds[i + 1]), bounds[i + 1]
        if a < b:
            n = max(1, int(round((b - a) / 0.01)))   # rollout::Settings::timeStep = 1e-2 (the adaptive integrator's first step)
            t += [a + (b - a) * j / n for j in range(n)] + [b]
        else:
            t.append(b)
        if i < len(bounds) - 2:
            post.append(len(t))
    tags = [mode_at_time(ev, modes, a, k == len(t) - 1) for k, a in enumerate(t)]
    event_data = [mode_at_time(ev, modes, t[i - 1]) for i in post]
    return np.array(t), np.array(tags), post, np.array(event_data, float)


CASES = {
    # name: (eventTimes, modeSequence, updatedEventTimes, updatedModeSequence, period, (willTruncate, willSpread) or None)
    "no_matching_modes": ([0.6, 1.7], [0, 1, 2], [1.0, 1.1], [10, 11, 12], (0.0, 2.0), (True, False)),
    "partially_matching_modes": ([0.6, 1.7, 2.3], [0, 1, 2, 3], [0.9, 1.1, 2.1], [10, 1, 2, 30], (1.0, 2.5), (True, True)),
    "final_time_is_the_same_as_event_time_1": ([1.1, 1.3], [0, 1, 2], [1.1, 2.1], [0, 1, 2], (0.2, 2.1), (False, True)),
    "final_time_is_the_same_as_event_time_2": ([1.1, 2.1], [0, 1, 2], [1.1, 1.3], [0, 1, 2], (0.2, 2.1), (False, True)),
    "erase_trajectory": ([1.1, 1.3], [0, 1, 2], [1.1, 1.3], [0, 1, 3], (0.2, 2.1), (True, False)),
    "fully_matched_modes": ([1.1, 1.3], [0, 1, 2], [0.5, 2.1], [0, 1, 2], (0.0, 2.5), (False, True)),
    "out_range_event_to_in_range_at_back_1": ([0.6, 1.7], [0, 1, 2], [1.0, 1.1], [0, 1, 2], (0.0, 1.5), None),
    "out_range_event_to_in_range_at_back_2": ([1.0], [0, 1], [1.0, 2.0], [0, 1, 2], (0.7, 2.5), None),
    "in_range_event_to_out_range_at_back_1": ([1.0, 1.1], [0, 1, 2], [0.6, 1.7], [0, 1, 2], (0.0, 1.5), None),
    "in_range_event_to_out_range_at_back_2": ([1, 2, 3.1], [0, 1, 2, 4], [1, 3], [0, 1, 3], (0.5, 2.5), None),
    "in_range_event_to_out_range_in_front_1": ([1, 2], [0, 1, 2], [0.5, 1.6], [0, 1, 2], (0.7, 2.5), None),
    "in_range_event_to_out_range_in_front_2": ([1, 2], [0, 1, 2], [1.6], [1, 2], (0.7, 2.5), None),
    "out_range_event_to_in_range_in_front_1": ([2], [1, 2], [0.5, 1.6], [0, 1, 2], (0.7, 2.5), None),
    "out_range_event_to_in_range_in_front_2": ([0.5, 2], [0, 1, 2], [1, 1.6], [0, 1, 2], (0.7, 2.5), None),
    "overlap_forward": ([1, 1.5, 2], [0, 1, 2, 3], [1, 2.2, 2.3], [0, 1, 2, 3], (0.9, 2.5), None),
    "overlap_backward": ([1, 2.2, 3], [0, 1, 2, 3], [1, 1.5, 2], [0, 1, 2, 3], (0, 3.5), None),
    "anymal_test": ([1.00001, 1.40001, 1.80001, 2.20001, 2.60001], [15, 15, 7, 14, 11, 13], [1.51913, 1.91913, 2.31913, 2.71913, 3.11913],
                    [15, 7, 14, 11, 13, 7], (1.4, 2.4), None),
}


def spread(case):
    ev, modes, nev, nmodes, period, _ = case
    ev, nev = [float(e) for e in ev], [float(e) for e in nev]
    t, tags, post, event_data = rollout(ev, modes, period)
    out = orc.trajectory_spread(ev, modes, nev, nmodes, t, tags[:, None].astype(float), tags, event_data)
    return ev, nev, t, tags, post, out


def test_no_change():
    """TEST_F no_change (:259-279): identical schedules leave everything untouched"""
    ev, modes, period = [0.6, 1.7], [0, 1, 2], (0.0, 2.0)
    t, tags, post, event_data = rollout(ev, modes, period)
    out = orc.trajectory_spread(ev, modes, ev, modes, t, tags[:, None].astype(float), tags, event_data)
    assert not out["will_truncate"] and not out["will_spread"]
    assert np.array_equal(out["t"], t) and np.array_equal(out["tags"], tags) and out["post_event_indices"] == post
    assert np.array_equal(out["event_data"], event_data)


@pytest.mark.parametrize("name", sorted(CASES))
def test_reference_recipe(name):
    """checkResults (:178-245) on the oracle's output"""
    case = CASES[name]
    _, modes, _, nmodes, period, expect = case
    ev, nev, t, tags, post, out = spread(case)
    if expect is not None:
        assert (out["will_truncate"], out["will_spread"]) == expect
    st, stags, spost = out["t"], out["tags"], out["post_event_indices"]
    # post-event indices recomputed from the updated schedule and the spread time trajectory (:181-205)
    if len(st):
        first, last = bisect.bisect_right(nev, st[0]), bisect.bisect_right(nev, st[-1])
        want = []
        for i in range(first, last):
            if i == last - 1 and nev[i] == st[-1]:
                want.append(len(st) - 1)
            else:
                want.append(bisect.bisect_right(list(st), nev[i]))
        assert spost == want
        assert abs(st[0] - period[0]) < 1e-6                                    # the initial time does not change (:210-212)
    else:
        assert spost == []
    ref_i = bisect.bisect_left(nev, period[0])                                  # lookup::findIndexInTimeArray
    it = iter(spost + [None])
    nxt = next(it)
    for k in range(len(st)):
        if 0 < k < len(st) - 1:
            assert st[k - 1] < st[k], k                                         # strictly increasing except the last pair (:219-221)
        if nxt is not None and nxt == k + 1:
            assert st[k] == nev[ref_i]                                          # the pre-event sample sits on the new event time (:224-226)
            ref_i += 1
            nxt = next(it)
        assert stags[k] == mode_at_time(nev, nmodes, st[k], k == len(st) - 1), (k, st[k])   # every kept sample carries the new schedule's mode (:235-237)
    # event data: one entry per kept post-event index, and it is the mode before that event (:240-249)
    assert len(out["event_data"]) == len(spost)
    for k, i in enumerate(spost):
        assert mode_at_time(nev, nmodes, st[i - 1]) == int(out["event_data"][k])


@pytest.mark.parametrize("name", sorted(CASES))
def test_host_restatements_match_the_oracle_on_the_reference_recipes(name):
    """the product's C++ and Python trajectorySpread against the (now pinned) oracle: identical times, values and flags"""
    from wb_humanoid_mpc_b200 import host_lib, references

    case = CASES[name]
    _, modes, _, nmodes, _, _ = case
    ev, nev, t, tags, post, out = spread(case)
    x = tags[:, None].astype(float)
    ct, cx, cu, trunc, spr = host_lib.trajectory_spread(ev, modes, nev, nmodes, t, x, x)
    assert np.array_equal(ct, out["t"]) and np.array_equal(cx[:, 0], out["tags"]) and np.array_equal(cu, cx)
    assert (trunc, spr) == (out["will_truncate"], out["will_spread"])
    py = references.trajectory_spread(references.ModeSchedule(ev, modes), references.ModeSchedule(nev, nmodes), dict(t=t, x=x, u=x.copy()))
    assert np.array_equal(np.asarray(py["t"]), out["t"]) and np.array_equal(np.asarray(py["x"]).reshape(-1), out["tags"].astype(float))
    assert (py["will_truncate"], py["will_spread"]) == (out["will_truncate"], out["will_spread"])
    assert list(py["post_event_indices"]) == out["post_event_indices"]


@pytest.mark.parametrize("seed", range(12))
def test_host_restatements_match_the_oracle_on_random_schedule_changes(seed):
    from wb_humanoid_mpc_b200 import host_lib, references

    rng = np.random.default_rng(100 + seed)
    n_ev = int(rng.integers(1, 6))
    ev = np.sort(rng.uniform(0.2, 2.8, n_ev)).tolist()
    modes = [int(m) for m in rng.integers(0, 4, n_ev + 1)]
    k = int(rng.integers(0, 3))
    nev = np.sort(np.clip(np.array(ev) + rng.uniform(-0.3, 0.3, n_ev), 0.05, 2.95)).tolist()
    nmodes = list(modes)
    if k == 1:      # the new schedule drops its first mode
        nev, nmodes = nev[1:], nmodes[1:]
    elif k == 2:    # the tail changes
        nmodes[-1] = 9
    t0, tf = float(rng.uniform(0.0, 0.5)), float(rng.uniform(2.0, 3.0))
    t, tags, post, event_data = rollout(ev, modes, (t0, tf))
    x = tags[:, None].astype(float)
    out = orc.trajectory_spread(ev, modes, nev, nmodes, t, x, tags, event_data)
    ct, cx, cu, trunc, spr = host_lib.trajectory_spread(ev, modes, nev, nmodes, t, x, x)
    assert np.array_equal(ct, out["t"]) and np.array_equal(cx[:, 0], out["tags"])
    assert (trunc, spr) == (out["will_truncate"], out["will_spread"])
    py = references.trajectory_spread(references.ModeSchedule(ev, modes), references.ModeSchedule(nev, nmodes), dict(t=t, x=x, u=x.copy()))
    assert np.array_equal(np.asarray(py["t"]), out["t"]) and np.array_equal(np.asarray(py["x"]).reshape(-1), out["tags"].astype(float))
