// examples/performance_eval.cpp -- the reference's PerformanceEval Monte Carlo (test/PerformanceEval.cpp) on the GPU:
// the same mt19937(0) trial stream and sampling ranges (:59-76, :130-148), the same acceptance filter
// (CheckInputFeasibility(5, 30, 20, 0.002) == InputFeasible, :163-173), the same check (one random sphere per accepted
// trial, minTimeSection 0.002, :180-182) -- but trials are drawn in blocks and every stage is one batched call through
// the C++ facade. Prints the three percentage lines of the reference's report in its format
// (data/PerformanceEval.txt:23-25 holds the reference's: 96.0038 / 3.99612 / 8e-05).
//
//   performance_eval [numTrajToGenerate = 10000000]
#include <chrono>
#include <iostream>
#include <random>
#include <vector>

#include "rqcd/facade.hpp"

using CommonMath::Vec3;

int main(int argc, char** argv) {
  const long long numTrajToGenerate = argc > 1 ? atoll(argv[1]) : 10000000LL;
  const double fmin = 5, fmax = 30, wmax = 20, minTimeSec = 0.002;
  std::mt19937 gen(0);
  std::uniform_real_distribution<> randPos(-4.0, 4.0), randVel(-4.0, 4.0), randAcc(-4.0, 4.0), randTime(0.2, 4.0),
      randRadius(0.1, 1.5);
  try {
    rqcd::Context ctx(0);
    rqcd::Batch batch(ctx);
    const long long block = 1 << 21;  // trials drawn per round
    std::vector<double> bc((size_t)RQCD_BC_ROWS * block), sph((size_t)4 * block);
    std::vector<double> abc, asph;  // accepted trials of this round, compacted
    unsigned long long counts[3] = {0, 0, 0};
    long long accepted = 0, trials = 0;
    double gpu_s = 0;
    while (accepted < numTrajToGenerate) {
      for (long long i = 0; i < block; i++) {
        // one trial = 20 draws in the driver's order; a constructor call `Vec3 v(r(gen), r(gen), r(gen))` evaluates
        // its arguments right to left with this compiler, so the first draw of each triple is the z component
        auto draw3 = [&](std::uniform_real_distribution<>& d, int row) {
          const double z = d(gen), y = d(gen), x = d(gen);
          bc[(size_t)(row + 0) * block + i] = x, bc[(size_t)(row + 1) * block + i] = y, bc[(size_t)(row + 2) * block + i] = z;
        };
        for (int k = 0; k < 3; k++) bc[(size_t)(RQCD_BC_P0 + k) * block + i] = 0.0;
        draw3(randVel, RQCD_BC_V0);
        draw3(randAcc, RQCD_BC_A0);
        draw3(randPos, RQCD_BC_PF);
        draw3(randVel, RQCD_BC_VF);
        draw3(randAcc, RQCD_BC_AF);
        bc[(size_t)RQCD_BC_TF * block + i] = randTime(gen);
        const double oz = randPos(gen), oy = randPos(gen), ox = randPos(gen);
        sph[i] = ox, sph[block + i] = oy, sph[2 * block + i] = oz;
        sph[3 * block + i] = randRadius(gen);
      }
      trials += block;
      auto t0 = std::chrono::steady_clock::now();
      std::vector<uint8_t> feas = batch.InputFeasibility(bc.data(), block, Vec3(0, 0, -9.81), fmin, fmax, wmax, minTimeSec);
      // keep the accepted trials, in order, up to the requested number
      long long m = 0;
      for (long long i = 0; i < block && accepted + m < numTrajToGenerate; i++) m += feas[i] == RQCD_INPUT_FEASIBLE;
      abc.assign((size_t)RQCD_BC_ROWS * m, 0.0);
      asph.assign((size_t)4 * m, 0.0);
      long long j = 0;
      for (long long i = 0; i < block && j < m; i++) {
        if (feas[i] != RQCD_INPUT_FEASIBLE) continue;
        for (int r = 0; r < RQCD_BC_ROWS; r++) abc[(size_t)r * m + j] = bc[(size_t)r * block + i];
        for (int r = 0; r < 4; r++) asph[(size_t)r * m + j] = sph[(size_t)r * block + i];
        j++;
      }
      rqcd::Batch::Result res = batch.GenerateAndCheckPairwise(abc.data(), asph.data(), m, minTimeSec);
      gpu_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      for (int v = 0; v < 3; v++) counts[v] += res.summary.traj_counts[v];
      accepted += m;
    }
    std::cout << "Done." << std::endl << std::endl;
    std::cout << "Trials drawn = " << trials << ", accepted = " << accepted << ", GPU calls [s] = " << gpu_s << std::endl;
    std::cout << "Percent of trajectories that were collision-free  = " << 100.0 * double(counts[0]) / numTrajToGenerate << std::endl;
    std::cout << "Percent of trajectories with collision  = " << 100.0 * double(counts[1]) / numTrajToGenerate << std::endl;
    std::cout << "Percent of trajectories with collision indeterminable  = " << 100.0 * double(counts[2]) / numTrajToGenerate
              << std::endl;
    return 0;
  } catch (const rqcd::Error& e) {
    std::cerr << e.what() << std::endl;
    return e.status() == RQCD_ERR_NO_DEVICE ? 3 : 2;
  }
}
