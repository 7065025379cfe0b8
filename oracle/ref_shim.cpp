/*
 * ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * A C-callable wrapper around the UNMODIFIED reference classes.  It is compiled together with the
 * reference's own sources, read from where they lie (REF=/root/reference, see oracle/Makefile), into
 * oracle/_ref/libref.so.  Nothing from the reference is copied into this repository: this file only
 * *calls* RapidTrajectoryGenerator / CollisionChecker / Sphere / RectPrism / Quartic through their
 * public headers, the way test/PerformanceEval.cpp and test/ForestPerformanceEval.cpp do.
 *
 * Uses: (1) validating oracle/oracle.c bit-for-bit, (2) generating tests/golden/ fixtures,
 * (3) the CPU baseline of bench.py (cpu_baseline.kind = "reference", --impl reference).
 * The product library never links it.
 */
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <random>
#include <thread>
#include <vector>

#include "CommonMath/ConvexObj.hpp"
#include "CommonMath/RectPrism.hpp"
#include "CommonMath/Sphere.hpp"
#include "RapidCollisionDetection/CollisionChecker.hpp"
#include "RapidQuadcopterTrajectories/RapidTrajectoryGenerator.hpp"

#include "../include/rqcd.h"

using namespace CommonMath;
using namespace RapidCollisionChecker;
using namespace RapidQuadrocopterTrajectoryGenerator;

namespace {

std::shared_ptr<ConvexObj> make_shape(const rqcd_obstacle& o) {
  Vec3 c(o.center[0], o.center[1], o.center[2]);
  if (o.type == RQCD_SPHERE) return std::make_shared<Sphere>(c, o.radius);
  return std::make_shared<RectPrism>(c, Vec3(o.side[0], o.side[1], o.side[2]),
                                     Rotation(o.quat[0], o.quat[1], o.quat[2], o.quat[3]));
}

Trajectory make_traj(const double* c, double t0, double t1) {
  return Trajectory(Vec3(c[0], c[1], c[2]), Vec3(c[3], c[4], c[5]), Vec3(c[6], c[7], c[8]),
                    Vec3(c[9], c[10], c[11]), Vec3(c[12], c[13], c[14]), Vec3(c[15], c[16], c[17]), t0, t1);
}

RapidTrajectoryGenerator make_gen(const rqcd_bc_soa* in, int64_t i, const double grav[3]) {
  const double* d = in->data;
  const int64_t s = in->stride;
  RapidTrajectoryGenerator g(Vec3(d[0 * s + i], d[1 * s + i], d[2 * s + i]), Vec3(d[3 * s + i], d[4 * s + i], d[5 * s + i]),
                             Vec3(d[6 * s + i], d[7 * s + i], d[8 * s + i]), Vec3(grav[0], grav[1], grav[2]));
  return g;
}

void set_goals(RapidTrajectoryGenerator& g, const rqcd_bc_soa* in, int64_t i) {
  const double* d = in->data;
  const int64_t s = in->stride;
  for (unsigned k = 0; k < 3; k++) {
    if (in->goal_mask & RQCD_GOAL_POS(k)) g.SetGoalPositionInAxis(k, d[(RQCD_BC_PF + k) * s + i]);
    if (in->goal_mask & RQCD_GOAL_VEL(k)) g.SetGoalVelocityInAxis(k, d[(RQCD_BC_VF + k) * s + i]);
    if (in->goal_mask & RQCD_GOAL_ACC(k)) g.SetGoalAccelerationInAxis(k, d[(RQCD_BC_AF + k) * s + i]);
  }
}

int check_pair(Trajectory& traj, const rqcd_obstacle& o, const std::shared_ptr<ConvexObj>& shape, double minT) {
  if (!o.moving) {
    CollisionChecker checker(traj);
    return (int)checker.CollisionCheck(shape, minT);
  }
  Trajectory obsTraj = make_traj(o.motion, o.motion_t_start, o.motion_t_end);
  double s = std::max(traj.GetStartTime(), obsTraj.GetStartTime());
  double e = std::min(traj.GetEndTime(), obsTraj.GetEndTime());
  if (!(s <= e)) return RQCD_NO_COLLISION; /* the reference would assert (Trajectory.hpp:68) */
  Trajectory rel = traj - obsTraj;
  CollisionChecker checker(rel);
  return (int)checker.CollisionCheck(shape, minT);
}

void summary_init(rqcd_summary* s) {
  std::memset(s, 0, sizeof *s);
  s->best_cost = std::numeric_limits<double>::infinity();
  s->best_index = -1;
}

struct Job {
  int mode;  // 0 generate_check, 1 check, 2 pairwise
  const rqcd_bc_soa* bc;
  const rqcd_coeff_soa* cf;
  const double* cost_in;
  const double* spheres;
  int64_t sphere_stride;
  int64_t n;
  const rqcd_obstacle* obs;
  int32_t m;
  double minT;
  uint32_t flags;
  int64_t index_base;
  rqcd_out* out;
};

void run_range(const Job& j, int64_t lo, int64_t hi, rqcd_summary* sum) {
  const double grav[3] = {0, 0, -9.81};
  summary_init(sum);
  std::vector<std::shared_ptr<ConvexObj>> shapes;
  for (int32_t k = 0; k < j.m; k++) shapes.push_back(make_shape(j.obs[k]));
  rqcd_out* out = j.out;
  for (int64_t i = lo; i < hi; i++) {
    double cost = std::numeric_limits<double>::quiet_NaN();
    double c[18];
    double t0 = 0, t1 = 0;
    if (j.mode == 1) {
      for (int k = 0; k < 18; k++) c[k] = j.cf->coeffs[k * j.cf->stride + i];
      t0 = j.cf->t_start ? j.cf->t_start[i] : 0.0;
      t1 = j.cf->t_end[i];
      if (j.cost_in) cost = j.cost_in[i];
    } else {
      RapidTrajectoryGenerator g = make_gen(j.bc, i, grav);
      set_goals(g, j.bc, i);
      g.Generate(j.bc->data[RQCD_BC_TF * j.bc->stride + i]);
      cost = g.GetCost();
      Trajectory t = g.GetTrajectory();
      for (int k = 0; k < 6; k++) {
        Vec3 v = t[k];
        c[3 * k] = v.x;
        c[3 * k + 1] = v.y;
        c[3 * k + 2] = v.z;
      }
      t0 = t.GetStartTime();
      t1 = t.GetEndTime();
      if (out->coeffs)
        for (int k = 0; k < 18; k++) out->coeffs[k * out->coeff_stride + i] = c[k];
      if (out->cost) out->cost[i] = cost;
    }
    Trajectory traj = make_traj(c, t0, t1);
    int verdict = RQCD_NO_COLLISION, first = -1;
    if (j.mode == 2) {
      std::shared_ptr<ConvexObj> sp = std::make_shared<Sphere>(
          Vec3(j.spheres[i], j.spheres[j.sphere_stride + i], j.spheres[2 * j.sphere_stride + i]),
          j.spheres[3 * j.sphere_stride + i]);
      CollisionChecker checker(traj);
      verdict = (int)checker.CollisionCheck(sp, j.minT);
      if (verdict != RQCD_NO_COLLISION) first = 0;
      sum->pair_counts[verdict]++;
      sum->pairs_checked++;
    } else {
      for (int32_t k = 0; k < j.m; k++) {
        int v = check_pair(traj, j.obs[k], shapes[k], j.minT);
        sum->pair_counts[v]++;
        sum->pairs_checked++;
        if (out->verdict_matrix) out->verdict_matrix[i * j.m + k] = (uint8_t)v;
        if (v != RQCD_NO_COLLISION && first < 0) {
          verdict = v;
          first = k;
          if (j.flags & RQCD_F_EARLY_EXIT) break;
        }
      }
    }
    if (out->verdict) out->verdict[i] = (uint8_t)verdict;
    if (out->first_obstacle) out->first_obstacle[i] = first;
    sum->traj_counts[verdict]++;
    if (verdict == RQCD_NO_COLLISION && cost < sum->best_cost) {
      sum->best_cost = cost;
      sum->best_index = j.index_base + i;
    }
  }
}

void run_job(Job& j, int nthreads) {
  rqcd_out dummy;
  std::memset(&dummy, 0, sizeof dummy);
  if (!j.out) j.out = &dummy;
  if (nthreads < 1) nthreads = 1;
  if ((int64_t)nthreads > j.n) nthreads = j.n > 0 ? (int)j.n : 1;
  std::vector<rqcd_summary> sums(nthreads);
  if (nthreads == 1) {
    run_range(j, 0, j.n, &sums[0]);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++)
      th.emplace_back([&j, &sums, t, nthreads]() { run_range(j, j.n * t / nthreads, j.n * (t + 1) / nthreads, &sums[t]); });
    for (auto& t : th) t.join();
  }
  if (j.out->summary) {
    rqcd_summary* m = j.out->summary;
    summary_init(m);
    for (int t = 0; t < nthreads; t++) {
      for (int v = 0; v < 4; v++) {
        m->traj_counts[v] += sums[t].traj_counts[v];
        m->pair_counts[v] += sums[t].pair_counts[v];
      }
      m->pairs_checked += sums[t].pairs_checked;
      if (sums[t].best_index >= 0 && sums[t].best_cost < m->best_cost) { /* shards ascend, so ties keep the lowest index */
        m->best_cost = sums[t].best_cost;
        m->best_index = sums[t].best_index;
      }
    }
  }
}

}  // namespace

extern "C" {

int ref_solve_p3(double a, double b, double c, double* x) { return (int)Quartic::solveP3(a, b, c, x); }
int ref_solve_quartic(double a, double b, double c, double d, double* root) {
  return (int)Quartic::solve_quartic(a, b, c, d, root);
}

void ref_solve_cubic_batch(const double* coef, int64_t n, double* roots, int32_t* count) {
  for (int64_t i = 0; i < n; i++) {
    double x[3];
    count[i] = (int32_t)Quartic::solveP3(coef[i], coef[n + i], coef[2 * n + i], x);
    for (int k = 0; k < 3; k++) roots[k * n + i] = x[k];
  }
}

void ref_solve_quartic_batch(const double* coef, int64_t n, double* roots, int32_t* count) {
  for (int64_t i = 0; i < n; i++) {
    double nan = std::numeric_limits<double>::quiet_NaN();
    double x[4] = {nan, nan, nan, nan};
    count[i] = (int32_t)Quartic::solve_quartic(coef[i], coef[n + i], coef[2 * n + i], coef[3 * n + i], x);
    for (int k = 0; k < 4; k++) roots[k * n + i] = x[k];
  }
}

int ref_check_one(const double coef[18], double t_start, double t_end, const rqcd_obstacle* obs, double minT) {
  Trajectory traj = make_traj(coef, t_start, t_end);
  return check_pair(traj, *obs, make_shape(*obs), minT);
}

void ref_rotation_matrix(const double quat[4], double R[9]) {
  Rotation(quat[0], quat[1], quat[2], quat[3]).GetRotationMatrix(R);
}

void ref_traj_value(const double coef[18], double t, double out[3]) {
  Vec3 v = make_traj(coef, t, t).GetValue(t);
  out[0] = v.x;
  out[1] = v.y;
  out[2] = v.z;
}

void ref_generate(const rqcd_bc_soa* in, int64_t n, rqcd_out* out, double* abg) {
  const double grav[3] = {0, 0, -9.81};
  for (int64_t i = 0; i < n; i++) {
    RapidTrajectoryGenerator g = make_gen(in, i, grav);
    set_goals(g, in, i);
    g.Generate(in->data[RQCD_BC_TF * in->stride + i]);
    if (out && out->cost) out->cost[i] = g.GetCost();
    if (out && out->coeffs) {
      Trajectory t = g.GetTrajectory();
      for (int k = 0; k < 6; k++) {
        Vec3 v = t[k];
        out->coeffs[(3 * k) * out->coeff_stride + i] = v.x;
        out->coeffs[(3 * k + 1) * out->coeff_stride + i] = v.y;
        out->coeffs[(3 * k + 2) * out->coeff_stride + i] = v.z;
      }
    }
    if (abg)
      for (int k = 0; k < 3; k++) {
        abg[(0 + k) * n + i] = g.GetAxisParamAlpha(k);
        abg[(3 + k) * n + i] = g.GetAxisParamBeta(k);
        abg[(6 + k) * n + i] = g.GetAxisParamGamma(k);
      }
  }
}

void ref_generate_check(const rqcd_bc_soa* in, int64_t n, const rqcd_obstacle* obs, int32_t m, double minT,
                        uint32_t flags, int64_t index_base, rqcd_out* out, int nthreads) {
  Job j{0, in, nullptr, nullptr, nullptr, 0, n, obs, m, minT, flags, index_base, out};
  run_job(j, nthreads);
}

void ref_check(const rqcd_coeff_soa* in, const double* cost, int64_t n, const rqcd_obstacle* obs, int32_t m,
               double minT, uint32_t flags, int64_t index_base, rqcd_out* out, int nthreads) {
  Job j{1, nullptr, in, cost, nullptr, 0, n, obs, m, minT, flags, index_base, out};
  run_job(j, nthreads);
}

void ref_generate_check_pairwise(const rqcd_bc_soa* in, const double* spheres, int64_t sphere_stride, int64_t n,
                                 double minT, int64_t index_base, rqcd_out* out, int nthreads) {
  Job j{2, in, nullptr, nullptr, spheres, sphere_stride, n, nullptr, 0, minT, 0, index_base, out};
  run_job(j, nthreads);
}

/* CheckInputFeasibility per trajectory.  Consecutive trajectories with equal group[i] (same initial state
 * required) reuse ONE generator object, as inside one Forest batch, so SingleAxisTrajectory's cached
 * _accPeakTimes persist across them.  group == NULL: a fresh object per trajectory. */
void ref_input_feasibility(const rqcd_bc_soa* in, int64_t n, const double gravity[3], double fmin, double fmax,
                           double wmax, double minT, const int64_t* group, uint8_t* result) {
  std::unique_ptr<RapidTrajectoryGenerator> g;
  for (int64_t i = 0; i < n; i++) {
    if (!g || !group || group[i] != group[i - 1]) g.reset(new RapidTrajectoryGenerator(make_gen(in, i, gravity)));
    set_goals(*g, in, i);
    g->Generate(in->data[RQCD_BC_TF * in->stride + i]);
    result[i] = (uint8_t)g->CheckInputFeasibility(fmin, fmax, wmax, minT);
  }
}

/* The PerformanceEval trial loop (test/PerformanceEval.cpp:59-76, 117-210) without its timers: same
 * engine, same distributions, same draw expressions, same acceptance filter.  Fills the accepted trials'
 * boundary conditions, spheres and verdicts; returns the total number of trials drawn. */
int64_t ref_c1_run(int64_t n_accept, int filter, double* bc, int64_t stride, double* spheres, int64_t sphere_stride,
                   uint8_t* verdict) {
  std::mt19937 gen(0);
  std::uniform_real_distribution<> randPos(-4.0, 4.0), randVel(-4.0, 4.0), randAcc(-4.0, 4.0), randTime(0.2, 4.0),
      randRadius(0.1, 1.5);
  Vec3 gravity = Vec3(0, 0, -9.81);
  Vec3 pos0(0, 0, 0);
  int64_t trials = 0;
  for (int64_t i = 0; i < n_accept; i++) {
    trials++;
    Vec3 vel0(randVel(gen), randVel(gen), randVel(gen));
    Vec3 acc0(randAcc(gen), randAcc(gen), randAcc(gen));
    Vec3 posf(randPos(gen), randPos(gen), randPos(gen));
    Vec3 velf(randVel(gen), randVel(gen), randVel(gen));
    Vec3 accf(randAcc(gen), randAcc(gen), randAcc(gen));
    double Tf = randTime(gen);
    RapidTrajectoryGenerator traj(pos0, vel0, acc0, gravity);
    traj.SetGoalPosition(posf);
    traj.SetGoalVelocity(velf);
    traj.SetGoalAcceleration(accf);
    Vec3 obsPos(randPos(gen), randPos(gen), randPos(gen));
    double radius = randRadius(gen);
    std::shared_ptr<ConvexObj> obstacle = std::make_shared<Sphere>(obsPos, radius);
    traj.Generate(Tf);
    if (filter) {
      RapidTrajectoryGenerator::InputFeasibilityResult inputFeas = traj.CheckInputFeasibility(5, 30, 20, 0.002);
      if (inputFeas != RapidTrajectoryGenerator::InputFeasible) {
        i--;
        continue;
      }
    }
    CollisionChecker checker(traj.GetTrajectory());
    CollisionChecker::CollisionResult stateFeas = checker.CollisionCheck(obstacle, 0.002);
    for (int k = 0; k < 3; k++) {
      bc[(RQCD_BC_P0 + k) * stride + i] = pos0[k];
      bc[(RQCD_BC_V0 + k) * stride + i] = vel0[k];
      bc[(RQCD_BC_A0 + k) * stride + i] = acc0[k];
      bc[(RQCD_BC_PF + k) * stride + i] = posf[k];
      bc[(RQCD_BC_VF + k) * stride + i] = velf[k];
      bc[(RQCD_BC_AF + k) * stride + i] = accf[k];
      spheres[k * sphere_stride + i] = obsPos[k];
    }
    bc[RQCD_BC_TF * stride + i] = Tf;
    spheres[3 * sphere_stride + i] = radius;
    if (verdict) verdict[i] = (uint8_t)stateFeas;
  }
  return trials;
}

/* The five Forest prisms (test/ForestPerformanceEval.cpp:130-136) as ABI records. */
void ref_forest_obstacles(rqcd_obstacle out[5]) {
  Vec3 objPos[5] = {Vec3(-1.75, 1.5, 0), Vec3(0.5, -1.5, 0), Vec3(1.5, 0.5, 0), Vec3(-1.0, -1.0, 0), Vec3(0.0, 0.8, -0.3)};
  Rotation objRot[5] = {Rotation::Identity(), Rotation::Identity(), Rotation::Identity(),
                        Rotation::FromAxisAngle(Vec3(1, 0, 0).GetUnitVector(), 45 * M_PI / 180),
                        Rotation::FromAxisAngle(Vec3(1, 0, 0).GetUnitVector(), -45 * M_PI / 180)};
  std::memset(out, 0, sizeof(rqcd_obstacle) * 5);
  for (int i = 0; i < 5; i++) {
    out[i].type = RQCD_RECT_PRISM;
    for (int k = 0; k < 3; k++) out[i].center[k] = objPos[i][k];
    out[i].side[0] = 0.5;
    out[i].side[1] = 0.5;
    out[i].side[2] = 5;
    for (unsigned k = 0; k < 4; k++) out[i].quat[k] = objRot[i][k];
  }
}

/* The ForestPerformanceEval batch loop (test/ForestPerformanceEval.cpp:77-102, 167-231) without timers:
 * one generator object per batch (so the stale _accPeakTimes cache is reproduced), early exit over the
 * five prisms.  Fills boundary conditions, inputFeas and stateFeas per candidate. */
void ref_c2_run(int64_t n_batches, int64_t per_batch, double* bc, int64_t stride, uint8_t* input_feas,
                uint8_t* verdict, int64_t* pairs_executed) {
  std::mt19937 gen(0);
  std::uniform_real_distribution<> randPos(-2.5, 2.5), randVelYZ(-2.0, 2.0), randVelX(2.0, 8.0), randAccYZ(-2.0, 2.0),
      randAccX(4.0, 10.0), randTime(0.5, 2.0);
  Vec3 gravity = Vec3(0, 0, -9.81);
  Vec3 pos0(-2.5, 0, 0);
  Vec3 velf(0, 0, 0);
  Vec3 accf(0, 0, 0);
  rqcd_obstacle recs[5];
  ref_forest_obstacles(recs);
  std::vector<std::shared_ptr<ConvexObj>> obstacles;
  for (int i = 0; i < 5; i++) obstacles.push_back(make_shape(recs[i]));
  int64_t idx = 0, pairs = 0;
  for (int64_t simNum = 0; simNum < n_batches; simNum++) {
    Vec3 vel0(randVelX(gen), randVelYZ(gen), randVelYZ(gen));
    Vec3 acc0(randAccX(gen), randAccYZ(gen), randAccYZ(gen));
    RapidTrajectoryGenerator traj(pos0, vel0, acc0, gravity);
    traj.SetGoalVelocity(velf);
    traj.SetGoalAcceleration(accf);
    for (int64_t trajNum = 0; trajNum < per_batch; trajNum++, idx++) {
      Vec3 posf(randPos(gen), randPos(gen), randPos(gen));
      traj.SetGoalPosition(posf);
      double Tf = randTime(gen);
      traj.Generate(Tf);
      RapidTrajectoryGenerator::InputFeasibilityResult inputFeas = traj.CheckInputFeasibility(5, 30, 20, 0.002);
      CollisionChecker::CollisionResult stateFeas = CollisionChecker::NoCollision;
      for (auto obs : obstacles) {
        CollisionChecker checker(traj.GetTrajectory());
        stateFeas = checker.CollisionCheck(obs, 0.002);
        pairs++;
        if (stateFeas != CollisionChecker::NoCollision) break;
      }
      for (int k = 0; k < 3; k++) {
        bc[(RQCD_BC_P0 + k) * stride + idx] = pos0[k];
        bc[(RQCD_BC_V0 + k) * stride + idx] = vel0[k];
        bc[(RQCD_BC_A0 + k) * stride + idx] = acc0[k];
        bc[(RQCD_BC_PF + k) * stride + idx] = posf[k];
        bc[(RQCD_BC_VF + k) * stride + idx] = velf[k];
        bc[(RQCD_BC_AF + k) * stride + idx] = accf[k];
      }
      bc[RQCD_BC_TF * stride + idx] = Tf;
      if (input_feas) input_feas[idx] = (uint8_t)inputFeas;
      if (verdict) verdict[idx] = (uint8_t)stateFeas;
    }
  }
  if (pairs_executed) *pairs_executed = pairs;
}

int ref_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
