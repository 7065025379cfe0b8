// tests/cpp/forest_facade.cpp -- the candidate loop of the reference's ForestPerformanceEval driver written against
// the C++ facade (include/rqcd/facade.hpp), i.e. the code a user of the reference ends up with after switching:
// same class names, but the per-candidate loop body becomes one batched call. Results are compared with the CPU
// oracle (oracle/liboracle.so, test infrastructure) on the same inputs.
//
// exit code: 0 parity ok, 1 mismatch, 3 no GPU (the facade has no CPU fallback).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <vector>

#include "rqcd/facade.hpp"
#include "../../oracle/oracle.h"

using namespace CommonMath;
using namespace RapidQuadrocopterTrajectoryGenerator;
using namespace RapidCollisionChecker;

int main(int argc, char** argv) {
  const int batches = argc > 1 ? atoi(argv[1]) : 20, perBatch = 100;
  const int64_t n = (int64_t)batches * perBatch;
  const double minT = 0.002;
  try {
    // ---- the five trees (test/ForestPerformanceEval.cpp:130-136)
    Vec3 objPos[5] = {Vec3(-1.75, 1.5, 0), Vec3(0.5, -1.5, 0), Vec3(1.5, 0.5, 0), Vec3(-1.0, -1.0, 0), Vec3(0.0, 0.8, -0.3)};
    Rotation objRot[5] = {Rotation::Identity(), Rotation::Identity(), Rotation::Identity(),
                          Rotation::FromAxisAngle(Vec3(1, 0, 0), 45 * M_PI / 180),
                          Rotation::FromAxisAngle(Vec3(1, 0, 0), -45 * M_PI / 180)};
    std::vector<std::shared_ptr<ConvexObj>> obstacles;
    for (int i = 0; i < 5; i++) obstacles.push_back(std::make_shared<RectPrism>(objPos[i], Vec3(0.5, 0.5, 5), objRot[i]));

    // ---- candidates: the driver's own mt19937(0) stream, drawn by the oracle's generator
    std::vector<double> bc((size_t)RQCD_BC_ROWS * n);
    orc_c2_inputs(batches, perBatch, bc.data(), n);

    // ---- batched: one call replaces the candidate loop (ForestPerformanceEval.cpp:186-225)
    rqcd::Context ctx(0);
    rqcd::Batch batch(ctx);
    batch.SetObstacles(obstacles);
    rqcd::Batch::Result got = batch.GenerateAndCheck(bc.data(), n, minT, RQCD_GOAL_ALL, /*earlyExit=*/true, 0,
                                                     /*wantCoeffs=*/true);
    // the driver also calls CheckInputFeasibility(5, 30, 20, 0.002) on its per-batch generator object (:203-209)
    std::vector<int64_t> group(n);
    for (int64_t i = 0; i < n; i++) group[i] = i / perBatch;
    std::vector<uint8_t> inputFeas = batch.InputFeasibility(bc.data(), n, Vec3(0, 0, -9.81), 5, 30, 20, 0.002, group.data());
    if (argc > 2) {  // batch 0 in the driver's file formats (ForestPerformanceEval.cpp:139-147, 255-267)
      std::ofstream csv(std::string(argv[2]) + "/ForestPerfEval.csv");
      rqcd::WriteForestBatchCsv(csv, perBatch, got.coeffs.data(), n, bc.data() + (size_t)RQCD_BC_TF * n, inputFeas.data(),
                                got.verdict.data());
      std::ofstream trees(std::string(argv[2]) + "/ForestPerfEvaltreeParamsCSV.csv");
      rqcd::WriteTreeParamsCsv(trees, obstacles);
    }

    // ---- oracle on the same arrays
    rqcd_obstacle recs[5];
    for (int i = 0; i < 5; i++) recs[i] = obstacles[i]->Record();
    std::vector<uint8_t> want(n);
    std::vector<double> wantCost(n);
    rqcd_summary ws;
    rqcd_out wo;
    memset(&wo, 0, sizeof wo);
    wo.verdict = want.data();
    wo.cost = wantCost.data();
    wo.summary = &ws;
    rqcd_bc_soa in = {bc.data(), n, RQCD_GOAL_ALL, 0};
    orc_generate_check(&in, n, recs, 5, minT, RQCD_F_EARLY_EXIT, 0, &wo, 4, nullptr);
    int bad = 0;
    for (int64_t i = 0; i < n; i++) {
      if (got.verdict[i] != want[i]) bad++;
      if (std::fabs(got.cost[i] - wantCost[i]) > 1e-9 * std::fabs(wantCost[i])) bad++;
    }
    if (got.summary.pairs_checked != ws.pairs_checked || got.summary.best_index != ws.best_index) bad++;
    printf("batched: %lld candidates, %llu pair checks, %llu collision free, best %lld, mismatches %d\n", (long long)n,
           (unsigned long long)got.summary.pairs_checked, (unsigned long long)got.summary.traj_counts[0],
           (long long)got.summary.best_index, bad);

    // ---- single objects, exactly the reference's calling sequence, for the first candidates of batch 0
    int badSingle = 0;
    for (int i = 0; i < 8; i++) {
      auto col = [&](int row) { return Vec3(bc[(row + 0) * n + i], bc[(row + 1) * n + i], bc[(row + 2) * n + i]); };
      RapidTrajectoryGenerator traj(col(RQCD_BC_P0), col(RQCD_BC_V0), col(RQCD_BC_A0), Vec3(0, 0, -9.81));
      traj.SetGoalPosition(col(RQCD_BC_PF));
      traj.SetGoalVelocity(Vec3(0, 0, 0));
      traj.SetGoalAcceleration(Vec3(0, 0, 0));
      traj.Generate(bc[RQCD_BC_TF * n + i]);
      CollisionChecker::CollisionResult stateFeas = CollisionChecker::NoCollision;
      for (auto obs : obstacles) {
        CollisionChecker checker(traj.GetTrajectory());
        stateFeas = checker.CollisionCheck(obs, minT);
        if (stateFeas != CollisionChecker::NoCollision) break;
      }
      if ((int)stateFeas != want[i]) badSingle++;
      if (std::fabs(traj.GetCost() - wantCost[i]) > 1e-9 * std::fabs(wantCost[i])) badSingle++;
    }
    printf("single-object calls: 8 candidates, mismatches %d (%s)\n", badSingle,
           CollisionChecker::GetCollisionResultName(CollisionChecker::NoCollision));
    return (bad || badSingle) ? 1 : 0;
  } catch (const rqcd::Error& e) {
    fprintf(stderr, "%s\n", e.what());
    return e.status() == RQCD_ERR_NO_DEVICE ? 3 : 2;
  }
}
