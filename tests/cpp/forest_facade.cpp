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
    std::vector<uint8_t> inputFeas_ = batch.InputFeasibility(bc.data(), n, Vec3(0, 0, -9.81), 5, 30, 20, 0.002, group.data());
    if (argc > 2) {  // batch 0 in the driver's file formats (ForestPerformanceEval.cpp:139-147, 255-267)
      std::ofstream csv(std::string(argv[2]) + "/ForestPerfEval.csv");
      rqcd::WriteForestBatchCsv(csv, perBatch, got.coeffs.data(), n, bc.data() + (size_t)RQCD_BC_TF * n, inputFeas_.data(),
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

    // ---- the same call from a multi-GPU C++ host (rqcd::Group = rqcd_group_*): every visible GPU, or two shards on the
    // one GPU there is; verdicts, costs and the MERGED summary (counts, best candidate) must equal the single-context call
    {
      int ndev = 0;
      {
        rqcd_ctx* probe = nullptr;
        while (ndev < 8 && rqcd_create(&probe, ndev) == RQCD_OK) {
          rqcd_destroy(probe);
          ndev++;
        }
      }
      std::vector<int32_t> devs;
      if (ndev >= 2) for (int d = 0; d < ndev; d++) devs.push_back(d);
      else devs = {0, 0};
      rqcd::Group group(devs);
      group.SetObstacles(obstacles);
      rqcd::Batch::Result gg = group.GenerateAndCheck(bc.data(), n, minT, RQCD_GOAL_ALL, /*earlyExit=*/true);
      int badGroup = 0;
      for (int64_t i = 0; i < n; i++)
        if (gg.verdict[i] != got.verdict[i] || gg.first_obstacle[i] != got.first_obstacle[i] || gg.cost[i] != got.cost[i]) badGroup++;
      if (memcmp(&gg.summary, &got.summary, sizeof(rqcd_summary)) != 0) badGroup++;
      printf("group of %d shards on %d device(s) (%s slots): mismatches %d\n", group.Size(), ndev >= 2 ? ndev : 1,
             group.PeerSlots() ? "peer-mapped" : "host-mapped", badGroup);
      bad += badGroup;
    }

    // ---- single objects: the reference's loop body UNCHANGED (ForestPerformanceEval.cpp:175-225) for the first two
    // batches -- one generator object per batch, SetGoalPosition / Generate / CheckInputFeasibility / GetTrajectory /
    // CollisionCheck per candidate. The per-object acceleration peak-time cache makes later candidates of a batch depend
    // on the first one (the `group` argument of the batched call); the facade object reproduces that.
    std::vector<uint8_t> wantFeas(n);
    const double grav3[3] = {0, 0, -9.81};
    orc_input_feasibility(&in, n, grav3, 5, 30, 20, 0.002, group.data(), wantFeas.data());
    std::vector<double> wantAbg((size_t)9 * n);
    {
      rqcd_out go;
      memset(&go, 0, sizeof go);
      orc_generate(&in, n, &go, wantAbg.data());
    }
    int badSingle = 0, singles = 0;
    const Vec3 velf(0, 0, 0), accf(0, 0, 0), gravity(0, 0, -9.81);
    for (int simNum = 0; simNum < 2 && simNum < batches; simNum++) {
      const int64_t b0 = (int64_t)simNum * perBatch;
      auto col = [&](int row, int64_t i) { return Vec3(bc[(row + 0) * n + i], bc[(row + 1) * n + i], bc[(row + 2) * n + i]); };
      RapidTrajectoryGenerator traj(col(RQCD_BC_P0, b0), col(RQCD_BC_V0, b0), col(RQCD_BC_A0, b0), gravity);
      traj.SetGoalVelocity(velf);
      traj.SetGoalAcceleration(accf);
      for (int trajNum = 0; trajNum < 24; trajNum++) {
        const int64_t i = b0 + trajNum;
        traj.SetGoalPosition(col(RQCD_BC_PF, i));
        const double Tf = bc[RQCD_BC_TF * n + i];
        traj.Generate(Tf);
        RapidTrajectoryGenerator::InputFeasibilityResult inputFeas = traj.CheckInputFeasibility(5, 30, 20, minT);
        CollisionChecker::CollisionResult stateFeas = CollisionChecker::NoCollision;
        for (auto obs : obstacles) {
          CollisionChecker checker(traj.GetTrajectory());
          stateFeas = checker.CollisionCheck(obs, minT);
          if (stateFeas != CollisionChecker::NoCollision) break;
        }
        singles++;
        if ((int)stateFeas != want[i]) badSingle++, fprintf(stderr, "candidate %lld: verdict %d, want %d\n", (long long)i, (int)stateFeas, (int)want[i]);
        if ((int)inputFeas != wantFeas[i] || (int)inputFeas != inputFeas_[i]) badSingle++, fprintf(stderr, "candidate %lld: input feasibility %d, oracle %d, batched %d\n", (long long)i, (int)inputFeas, (int)wantFeas[i], (int)inputFeas_[i]);
        if (std::fabs(traj.GetCost() - wantCost[i]) > 1e-9 * std::fabs(wantCost[i])) badSingle++;
        if (traj.GetFinalTime() != Tf) badSingle++;
        for (int ax = 0; ax < 3; ax++) {
          const double w[3] = {wantAbg[(0 + ax) * n + i], wantAbg[(3 + ax) * n + i], wantAbg[(6 + ax) * n + i]};
          const double g[3] = {traj.GetAxisParamAlpha(ax), traj.GetAxisParamBeta(ax), traj.GetAxisParamGamma(ax)};
          for (int q = 0; q < 3; q++)
            if (!(std::fabs(g[q] - w[q]) <= 1e-9 * std::fabs(w[q]))) badSingle++;
        }
        // CollisionCheck(Boundary) and CheckPositionFeasibility are the same plane test (CollisionChecker.cpp:25-68,
        // RapidTrajectoryGenerator.cpp:162-212): a floor below and a wall beside the forest
        const Boundary planes[2] = {{Vec3(0, 0, -1.0), Vec3(0, 0, 1)}, {Vec3(2.0, 0, 0), Vec3(-1, 0.2, 0)}};
        for (const Boundary& bnd : planes) {
          CollisionChecker checker(traj.GetTrajectory());
          const CollisionChecker::CollisionResult viaChecker = checker.CollisionCheck(bnd);
          const RapidTrajectoryGenerator::StateFeasibilityResult viaGen = traj.CheckPositionFeasibility(bnd.point, bnd.normal);
          const Trajectory t = traj.GetTrajectory();
          double k[18];
          for (int a = 0; a < 6; a++)
            for (int d = 0; d < 3; d++) k[3 * a + d] = t[a][d];
          const double pl[6] = {bnd.point.x, bnd.point.y, bnd.point.z, bnd.normal.x, bnd.normal.y, bnd.normal.z};
          const int wantPlane = orc_check_plane_one(k, t.GetStartTime(), t.GetEndTime(), pl);
          if ((int)viaChecker != wantPlane || (int)viaGen != wantPlane) badSingle++, fprintf(stderr, "candidate %lld: plane %d / %d, want %d\n", (long long)i, (int)viaChecker, (int)viaGen, wantPlane);
        }
      }
    }
    // ---- a translating obstacle through Trajectory::operator- (Trajectory.hpp:121-128) equals the MovingObj record of
    // the batched call: CollisionChecker(traj - obstacleTraj).CollisionCheck(shapeAtOrigin, minT)
    {
      auto col = [&](int row, int64_t i) { return Vec3(bc[(row + 0) * n + i], bc[(row + 1) * n + i], bc[(row + 2) * n + i]); };
      RapidTrajectoryGenerator obsGen(Vec3(0.5, -1.0, 0.2), Vec3(0.3, 0.8, 0), Vec3(0, 0, 0), gravity);
      obsGen.SetGoalPosition(Vec3(-0.5, 1.5, 0.0));
      obsGen.SetGoalVelocity(velf);
      obsGen.Generate(1.7);
      const Trajectory obsTraj = obsGen.GetTrajectory();
      auto shape = std::make_shared<RectPrism>(Vec3(0, 0, 0), Vec3(0.6, 0.4, 2.0), Rotation::FromRotationVector(Vec3(0.3, -0.2, 0.5)));
      std::vector<std::shared_ptr<ConvexObj>> moving = {std::make_shared<MovingObj>(shape, obsTraj)};
      rqcd::Context ctx2(0);  // its own obstacle table
      rqcd::Batch mb(ctx2);
      mb.SetObstacles(moving);
      const int64_t nm = 64;
      std::vector<double> sub((size_t)RQCD_BC_ROWS * nm);
      for (int r = 0; r < RQCD_BC_ROWS; r++)
        for (int64_t i = 0; i < nm; i++) sub[(size_t)r * nm + i] = bc[(size_t)r * n + i];
      rqcd::Batch::Result mr = mb.GenerateAndCheck(sub.data(), nm, minT);
      for (int64_t i = 0; i < nm; i++) {
        RapidTrajectoryGenerator traj(col(RQCD_BC_P0, i), col(RQCD_BC_V0, i), col(RQCD_BC_A0, i), gravity);
        traj.SetGoalPosition(col(RQCD_BC_PF, i));
        traj.SetGoalVelocity(velf);
        traj.SetGoalAcceleration(accf);
        traj.Generate(bc[RQCD_BC_TF * n + i]);
        CollisionChecker checker(traj.GetTrajectory() - obsTraj);
        singles++;
        if ((int)checker.CollisionCheck(shape, minT) != mr.verdict[i]) badSingle++, fprintf(stderr, "candidate %lld: operator- check differs from the MovingObj batch\n", (long long)i);
      }
      // the single-object calls above (each uploads a one-entry obstacle table) ran on the facade's private context:
      // the forest table of `batch` on `ctx` is untouched and the batched call repeats exactly
      rqcd::Batch::Result again = batch.GenerateAndCheck(bc.data(), n, minT, RQCD_GOAL_ALL, true);
      if (again.verdict != got.verdict || again.summary.pairs_checked != got.summary.pairs_checked) {
        badSingle++;
        fprintf(stderr, "the batch's obstacle table was disturbed\n");
      }
    }
    printf("single-object calls: %d candidates, mismatches %d (%s)\n", singles, badSingle,
           CollisionChecker::GetCollisionResultName(CollisionChecker::NoCollision));
    return (bad || badSingle) ? 1 : 0;
  } catch (const rqcd::Error& e) {
    fprintf(stderr, "%s\n", e.what());
    return e.status() == RQCD_ERR_NO_DEVICE ? 3 : 2;
  }
}
