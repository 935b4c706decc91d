// sqp::Logger CSV output for b200sqp::host::SqpSolver, column for column the format the reference writes when compiled with logging
// (lib/ocs2_ros2/ocs2_sqp/ocs2_sqp/src/SqpLogging.cpp:36-92, include/ocs2_sqp/SqpLogging.h:44-64), so that the reference's tooling
// (ReadSqpLog.py) reads the logs of a batch unchanged: one line per (instance, iteration); `problemNumber` = the instance index.
// The three time columns hold the DEVICE milliseconds of the whole batch's stage (SqpSolver::getBenchmarks) divided by the batch size: the
// per-instance amortised cost, which is what a batched solve has in place of a per-problem wall clock.
#pragma once
#include <cmath>
#include <iomanip>
#include <ostream>
#include <sstream>
#include <string>

#include "SqpSolver.hpp"

namespace b200sqp::host {

inline std::string toString(int stepType) {  // FilterLinesearch::StepType (FilterLinesearch.cpp:59-74)
  switch (stepType) {
    case 1: return "Constraint";
    case 2: return "Dual";
    case 3: return "Cost";
    case 4: return "Zero";
    default: return "Unknown";
  }
}
inline std::string convergenceToString(int c) {  // sqp::Convergence (SqpSolverStatus.h:60-74)
  switch (c) {
    case 1: return "Maximum number of iterations reached";
    case 2: return "Step size below minimum";
    case 3: return "Cost decrease and constraint satisfaction below tolerance";
    case 4: return "Primal update below tolerance";
    default: return "Not Converged";
  }
}
// FilterLinesearch::totalConstraintViolation (FilterLinesearch.h:63-65)
inline double totalConstraintViolation(const PerformanceIndex& p) { return std::sqrt(p.dynamicsViolationSSE + p.equalityConstraintsSSE); }

inline std::string logHeader() {
  return "problemNumber, time, iteration, linearQuadraticApproximationTime, solveQpTime, linesearchTime, baselinePerformanceIndex/merit, "
         "baselinePerformanceIndex/dynamicsViolationSSE, baselinePerformanceIndex/equalityConstraintsSSE, totalConstraintViolationBaseline, "
         "stepSize, stepType, dxNorm, duNorm, performanceAfterStep/merit, performanceAfterStep/dynamicsViolationSSE, "
         "performanceAfterStep/equalityConstraintsSSE, totalConstraintViolationAfterStep, convergence\n";
}

// iteration log of every instance of the last run(); `time` = the horizon start time passed to run()
inline void writeLog(std::ostream& stream, const SqpSolver& solver, double time) {
  const std::string d = ", ";
  const Benchmarks bm = solver.getBenchmarks();
  const double B = solver.batch();
  stream << std::setprecision(16);
  for (int b = 0; b < solver.batch(); ++b) {
    size_t it = 0;
    for (const StepInfo& s : solver.getIterationsLog(b)) {
      stream << b << d << time << d << it++ << d << bm.linearQuadraticApproximation / B << d << bm.solveQp / B << d << bm.linesearch / B << d
             << s.baseline.merit << d << s.baseline.dynamicsViolationSSE << d << s.baseline.equalityConstraintsSSE << d
             << totalConstraintViolation(s.baseline) << d << s.stepSize << d << toString(s.stepType) << d << s.dx_norm << d << s.du_norm << d
             << s.performanceAfterStep.merit << d << s.performanceAfterStep.dynamicsViolationSSE << d
             << s.performanceAfterStep.equalityConstraintsSSE << d << totalConstraintViolation(s.performanceAfterStep) << d
             << convergenceToString(s.convergence) << "\n";
    }
  }
}

}  // namespace b200sqp::host
