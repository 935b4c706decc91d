// Minimal C++ application on the host layer: a batch of whole-body MPC instances solved on one GPU through b200sqp::host::SqpSolver, the
// mirror of ocs2::SqpSolver.  No dependency beyond libb200sqp.so and the C++17 standard library.
//
//   g++ -std=c++17 -O2 -I. examples/host_batch.cpp -Lwb_humanoid_mpc_b200 -lb200sqp -pthread -Wl,-rpath,$PWD/wb_humanoid_mpc_b200 -o host_batch
//   ./host_batch wb_humanoid_mpc_b200/data/g1_wb_model.txt 64          (or data/g1_centroidal_model.txt)
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <random>

#include "wb_humanoid_mpc_b200/host/SqpLogging.hpp"
#include "wb_humanoid_mpc_b200/host/SqpSolver.hpp"

using namespace b200sqp::host;

int main(int argc, char** argv) {
  const std::string path = argc > 1 ? argv[1] : "wb_humanoid_mpc_b200/data/g1_wb_model.txt";
  const int batch = argc > 2 ? std::atoi(argv[2]) : 64;
  try {
    const HostModel model = loadModelFile(path);
    SqpSolver solver(model, model.sqpSettings, batch, /*device=*/0);
    std::mt19937_64 rng(1234);
    std::uniform_real_distribution<double> vx(-0.5, 1.0), yaw(-0.5, 0.5);
    std::vector<vector_t> x0(batch, model.initialState);
    const double t0 = 0.0, tf = t0 + model.timeHorizon;
    for (int b = 0; b < batch; ++b) {
      solver.getReferenceManager(b).setGait(b % 2 ? "walk" : "stance", t0, tf + model.timeHorizon);
      const std::array<double, 4> cmd{vx(rng), 0.0, model.defaultBaseHeight, yaw(rng)};
      solver.getReferenceManager(b).setTargetTrajectories(
          model.centroidal ? commandedVelocityToTargetTrajectoriesCentroidal(model, t0, x0[b], cmd, model.timeHorizon, centroidalBaseVelocity(model, x0[b]))
                           : commandedVelocityToTargetTrajectories(model, t0, x0[b], cmd, model.timeHorizon));
    }
    for (int cycle = 0; cycle < 3; ++cycle) {   // receding horizon with perfect tracking: the next initial state is the planned one
      const double t = t0 + cycle * model.dt;
      solver.run(t, x0, t + model.timeHorizon);
      for (int b = 0; b < batch; ++b) x0[b] = linearInterpolate(t + model.dt, solver.primalSolution(b).timeTrajectory_, solver.primalSolution(b).stateTrajectory_);
      const Benchmarks bm = solver.getBenchmarks();
      std::printf("cycle %d: LQ %.2f ms, QP %.2f ms, line search %.2f ms (device, whole batch); instance 0: step %.3f, merit %.4f\n", cycle,
                  bm.linearQuadraticApproximation, bm.solveQp, bm.linesearch, solver.getIterationsLog(0).back().stepSize,
                  solver.getIterationsLog(0).back().performanceAfterStep.merit);
    }
    std::ofstream csv("sqp_log.csv");
    csv << logHeader();
    writeLog(csv, solver, t0 + 2 * model.dt);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "host_batch: %s\n", e.what());
    return 1;
  }
  return 0;
}
