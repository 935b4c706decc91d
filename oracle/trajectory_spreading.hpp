// CPU oracle (TEST INFRASTRUCTURE ONLY -- never linked into the product): restatement of ocs2::TrajectorySpreading, the first thing
// SqpSolver::runImpl does to the previous primal solution (SqpSolver.cpp:211-213 -> trajectorySpread,
// ocs2_oc/include/ocs2_oc/trajectory_adjustment/TrajectorySpreadingHelperFunctions.h:124-143).
//
// Follows lib/ocs2_ros2/ocs2_oc/src/trajectory_adjustment/TrajectorySpreading.cpp:52-166 (set), :271-350 (computeSpreadingStrategy),
// :355-367 (adjustTimeTrajectory) and the templates of include/ocs2_oc/trajectory_adjustment/TrajectorySpreading.h:124-183
// (findPostEventIndices, extractEventsArray, adjustTrajectory).  Written as one pure function that returns a plan plus three appliers, so that
// it shares no code with the product's host restatements (host/references.hpp, references.py) it is used to pin.
// Pinned by the reference's own recipes: all 17 cases of ocs2_oc/test/trajectory_adjustment/TrajectorySpreadingTest.cpp with the
// property checks of its checkResults() and the Status flags it EXPECTs (tests/test_oracle_spreading.py).
#pragma once
#include <algorithm>
#include <cstddef>
#include <vector>

namespace oracle {

struct SpreadSchedule {
  std::vector<double> eventTimes;
  std::vector<int> modeSequence;  // eventTimes.size() + 1 entries
};

struct SpreadPlan {
  size_t eraseFrom = 0;                        // first erased sample (TrajectorySpreading.cpp:135-145)
  size_t keepEventsFirst = 0, keepEventsLast = 0;  // keepEventDataInInterval_ [first, last)
  struct Copy {
    size_t begin, end, from;                   // samples [begin, end) take the value the sample `from` had before spreading
  };
  std::vector<Copy> copies;
  std::vector<size_t> postEventIndices;        // updatedPostEventIndices_
  std::vector<double> postEventTimes;          // updatedMatchedEventTimes_
  bool willTruncate = false, willSpread = false;
};

namespace spreading_detail {
inline size_t firstAfter(const std::vector<double>& v, double t) {  // std::upper_bound index
  size_t i = 0;
  while (i < v.size() && !(t < v[i])) ++i;
  return i;
}
inline size_t firstNotBefore(const std::vector<double>& v, double t) {  // std::lower_bound index
  size_t i = 0;
  while (i < v.size() && v[i] < t) ++i;
  return i;
}
// TrajectorySpreading.h:124-134: index of the first sample after each event; an event exactly at the final time maps to the last sample
inline std::vector<size_t> postEventIndices(const std::vector<double>& events, const std::vector<double>& time) {
  std::vector<size_t> out;
  for (size_t i = 0; i < events.size(); ++i) {
    const bool lastAtEnd = (i + 1 == events.size()) && events[i] == time.back();
    out.push_back(lastAtEnd ? time.size() - 1 : firstAfter(time, events[i]));
  }
  return out;
}
}  // namespace spreading_detail

inline SpreadPlan spreadPlan(const SpreadSchedule& oldMs, const SpreadSchedule& newMs, const std::vector<double>& oldTime) {
  using namespace spreading_detail;
  SpreadPlan plan;
  const double tBegin = oldTime.front(), tEnd = oldTime.back();
  // modes the old solution contains / the new schedule needs over the same period (cpp:56-66)
  const int oldFirst = static_cast<int>(firstAfter(oldMs.eventTimes, tBegin)), oldLast = static_cast<int>(firstAfter(oldMs.eventTimes, tEnd));
  const int newFirst = static_cast<int>(firstAfter(newMs.eventTimes, tBegin)), newLast = static_cast<int>(firstAfter(newMs.eventTimes, tEnd));
  // longest common prefix of old[oldStart..oldLast] and new[newFirst..newLast], oldStart advanced until the prefix is non-empty (cpp:73-95)
  int oldStart = oldFirst, window = 0;
  for (; oldStart < static_cast<int>(oldMs.modeSequence.size()); ++oldStart) {
    window = 0;
    for (int a = oldStart, b = newFirst; a <= oldLast && b <= newLast && oldMs.modeSequence[a] == newMs.modeSequence[b]; ++a, ++b) ++window;
    if (window > 0) break;
  }
  std::vector<double> oldEv, newEv;  // pairs of (old event time, where it has to go)
  if (window > 0) {
    // phase 1 (cpp:104-116): the window - 1 events between matched modes
    for (int i = 0; i + 1 < window; ++i) {
      oldEv.push_back(oldMs.eventTimes[oldStart + i]);
      newEv.push_back(newMs.eventTimes[newFirst + i]);
    }
    plan.keepEventsFirst = static_cast<size_t>(oldStart - oldFirst);
    plan.keepEventsLast = plan.keepEventsFirst + static_cast<size_t>(window - 1);
    // phase 2 (cpp:118-129): the matched window starts later in the old schedule -> its triggering event moves before the initial time
    if (oldStart > oldFirst) {
      oldEv.insert(oldEv.begin(), oldMs.eventTimes[oldStart - 1]);
      newEv.insert(newEv.begin(), tBegin - 1e-4);
    }
  }
  // phase 3 (cpp:131-152): the event that ends the matched window, if it has to move forward (or beyond the end)
  const bool oldTailMatched = (oldStart + window - 1 == oldLast);
  const bool newTailMatched = (newFirst + window - 1 == newLast);
  if (window > 0 && !oldTailMatched) {  // (window == 0 erases everything; the reference indexes out of range there)
    const double oldEnd = oldMs.eventTimes[oldStart + window - 1];
    if (newTailMatched) {
      oldEv.push_back(oldEnd);
      newEv.push_back(tEnd + 1e-4);
    } else if (oldEnd < newMs.eventTimes[newFirst + window - 1]) {
      oldEv.push_back(oldEnd);
      newEv.push_back(newMs.eventTimes[newFirst + window - 1]);
    }
  }
  // truncation point (cpp:156-166)
  plan.eraseFrom = oldTime.size();
  if (window == 0) plan.eraseFrom = 0;
  else if (!newTailMatched) plan.eraseFrom = firstNotBefore(oldTime, newMs.eventTimes[newFirst + window - 1]);
  // spreading intervals (cpp:271-350)
  const std::vector<size_t> was = postEventIndices(oldEv, oldTime), goes = postEventIndices(newEv, oldTime);
  for (size_t j = 0; j < was.size(); ++j) {
    if (goes[j] < was[j]) {  // the event moves earlier: the post-event value fills [new, old)
      plan.copies.push_back({goes[j], std::min(was[j], plan.eraseFrom), was[j]});
    } else if (goes[j] > was[j]) {  // the event moves later: the pre-event value fills [old, new), never over the previous event's stretch
      const size_t begin = (j == 0) ? was[j] : std::max(was[j], goes[j - 1]);
      plan.copies.push_back({begin, goes[j], was[j] - 1});
    }
    if (goes[j] != 0 && goes[j] < plan.eraseFrom) {
      plan.postEventIndices.push_back(goes[j]);
      plan.postEventTimes.push_back(newEv[j]);
    }
  }
  plan.willTruncate = plan.eraseFrom < oldTime.size();
  plan.willSpread = !plan.copies.empty();
  return plan;
}

// TrajectorySpreading::adjustTrajectory (h:166-181): rows of `traj` are samples of `width` doubles; returns the new number of samples
inline size_t spreadApply(const SpreadPlan& plan, std::vector<double>& traj, size_t width) {
  traj.resize(plan.eraseFrom * width);
  std::vector<std::vector<double>> values;  // taken beforehand: spreading may overwrite its own sources
  for (const auto& c : plan.copies) values.emplace_back(traj.begin() + c.from * width, traj.begin() + (c.from + 1) * width);
  for (size_t i = 0; i < plan.copies.size(); ++i)
    for (size_t j = plan.copies[i].begin; j < plan.copies[i].end; ++j) std::copy(values[i].begin(), values[i].end(), traj.begin() + j * width);
  return plan.eraseFrom;
}

// TrajectorySpreading::adjustTimeTrajectory (cpp:355-367), eps = numeric_traits::weakEpsilon = 1e-9
inline void spreadApplyTime(const SpreadPlan& plan, std::vector<double>& time) {
  time.resize(plan.eraseFrom);
  for (size_t i = 0; i < plan.postEventIndices.size(); ++i) {
    const size_t k = plan.postEventIndices[i];
    time[k - 1] = plan.postEventTimes[i];
    time[k] = std::min(plan.postEventTimes[i] + 1e-9, time.back());
  }
}

}  // namespace oracle
