// examples/forest_performance_eval.cpp -- the reference's ForestPerformanceEval Monte Carlo
// (test/ForestPerformanceEval.cpp) on the GPU: the same mt19937(0) stream (:77-102, :175-193), the same five trees
// (:130-136), one generator object per batch for CheckInputFeasibility(5, 30, 20, 0.002) (:177-209) and the candidate
// loop with its break at the first obstacle hit (:216-225) -- all candidates of all batches in two batched calls.
// Prints the two percentage lines of the reference's report (data/ForestPerfEval.txt:28-29 holds 63.9533 / 60.381)
// and, with a directory argument, writes batch 0 in the driver's CSV formats.
//
//   forest_performance_eval [numSims = 10000] [output directory]
#include <chrono>
#include <cmath>
#include <fstream>
#include <iostream>
#include <memory>
#include <random>
#include <vector>

#include "rqcd/facade.hpp"

using namespace CommonMath;

int main(int argc, char** argv) {
  const long long numSims = argc > 1 ? atoll(argv[1]) : 10000, perSim = 100, n = numSims * perSim;
  const double minTimeSec = 0.002;
  std::mt19937 gen(0);
  std::uniform_real_distribution<> randPos(-2.5, 2.5), randVelYZ(-2.0, 2.0), randVelX(2.0, 8.0), randAccYZ(-2.0, 2.0),
      randAccX(4.0, 10.0), randTime(0.5, 2.0);
  try {
    Vec3 objPos[5] = {Vec3(-1.75, 1.5, 0), Vec3(0.5, -1.5, 0), Vec3(1.5, 0.5, 0), Vec3(-1.0, -1.0, 0), Vec3(0.0, 0.8, -0.3)};
    Rotation objRot[5] = {Rotation::Identity(), Rotation::Identity(), Rotation::Identity(),
                          Rotation::FromAxisAngle(Vec3(1, 0, 0), 45 * M_PI / 180),
                          Rotation::FromAxisAngle(Vec3(1, 0, 0), -45 * M_PI / 180)};
    std::vector<std::shared_ptr<ConvexObj>> obstacles;
    for (int i = 0; i < 5; i++) obstacles.push_back(std::make_shared<RectPrism>(objPos[i], Vec3(0.5, 0.5, 5), objRot[i]));

    std::vector<double> bc((size_t)RQCD_BC_ROWS * n, 0.0);
    std::vector<int64_t> group(n);
    auto at = [&](int row, long long i) -> double& { return bc[(size_t)row * n + i]; };
    long long i = 0;
    for (long long sim = 0; sim < numSims; sim++) {
      // `Vec3 vel0(randVelX(gen), randVelYZ(gen), randVelYZ(gen))`: arguments are evaluated right to left
      const double vz = randVelYZ(gen), vy = randVelYZ(gen), vx = randVelX(gen);
      const double az = randAccYZ(gen), ay = randAccYZ(gen), ax = randAccX(gen);
      for (long long t = 0; t < perSim; t++, i++) {
        const double pz = randPos(gen), py = randPos(gen), px = randPos(gen);
        at(RQCD_BC_P0, i) = -2.5;
        at(RQCD_BC_V0, i) = vx, at(RQCD_BC_V0 + 1, i) = vy, at(RQCD_BC_V0 + 2, i) = vz;
        at(RQCD_BC_A0, i) = ax, at(RQCD_BC_A0 + 1, i) = ay, at(RQCD_BC_A0 + 2, i) = az;
        at(RQCD_BC_PF, i) = px, at(RQCD_BC_PF + 1, i) = py, at(RQCD_BC_PF + 2, i) = pz;
        at(RQCD_BC_TF, i) = randTime(gen);
        group[i] = sim;
      }
    }
    rqcd::Context ctx(0);
    rqcd::Batch batch(ctx);
    batch.SetObstacles(obstacles);
    auto t0 = std::chrono::steady_clock::now();
    std::vector<uint8_t> inputFeas = batch.InputFeasibility(bc.data(), n, Vec3(0, 0, -9.81), 5, 30, 20, 0.002, group.data());
    rqcd::Batch::Result res = batch.GenerateAndCheck(bc.data(), n, minTimeSec, RQCD_GOAL_ALL, /*earlyExit=*/true, 0,
                                                     /*wantCoeffs=*/argc > 2);
    const double gpu_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    long long numInputFeasible = 0;
    for (uint8_t f : inputFeas) numInputFeasible += f == RQCD_INPUT_FEASIBLE;
    if (argc > 2) {
      std::ofstream csv(std::string(argv[2]) + "/ForestPerfEval.csv");
      rqcd::WriteForestBatchCsv(csv, perSim, res.coeffs.data(), n, bc.data() + (size_t)RQCD_BC_TF * n, inputFeas.data(),
                                res.verdict.data());
      std::ofstream trees(std::string(argv[2]) + "/ForestPerfEvaltreeParamsCSV.csv");
      rqcd::WriteTreeParamsCsv(trees, obstacles);
    }
    std::cout << "Done." << std::endl << std::endl;
    std::cout << "Candidates = " << n << ", executed trajectory-obstacle checks = " << res.summary.pairs_checked
              << ", GPU calls [s] = " << gpu_s << std::endl;
    std::cout << "Percent of trajectories that were input feasible = " << 100.0 * double(numInputFeasible) / perSim / numSims
              << std::endl;
    std::cout << "Percent of trajectories that were collision free = "
              << 100.0 * double(res.summary.traj_counts[0]) / perSim / numSims << std::endl;
    return 0;
  } catch (const rqcd::Error& e) {
    std::cerr << e.what() << std::endl;
    return e.status() == RQCD_ERR_NO_DEVICE ? 3 : 2;
  }
}
