// rqcd/facade.hpp -- C++ facade over the C ABI (rqcd.h) that keeps the reference's names for the hot path:
//
//   CommonMath::Vec3 / Rotation / Boundary / ConvexObj / Sphere / RectPrism / Trajectory   (data holders)
//   RapidQuadrocopterTrajectoryGenerator::RapidTrajectoryGenerator
//        ctor(x0, v0, a0, gravity), SetGoal{Position,Velocity,Acceleration}[InAxis], Reset, Generate, GetCost,
//        GetTrajectory, GetAxisParamAlpha/Beta/Gamma, GetFinalTime, CheckInputFeasibility, CheckPositionFeasibility
//                                         (reference: include/RapidQuadcopterTrajectories/RapidTrajectoryGenerator.hpp:63-241)
//   RapidCollisionChecker::CollisionChecker
//        ctor(Trajectory), CollisionCheck(Boundary), CollisionCheck(shared_ptr<ConvexObj>, minTimeSection)
//                                         (reference: include/RapidCollisionDetection/CollisionChecker.hpp:33-61)
//   rqcd::Batch -- the batched form both drivers' trial loops collapse to (PerformanceEval.cpp:140-182,
//                  ForestPerformanceEval.cpp:177-225): N candidates x M obstacles in one call.
//
// Every method that computes something is one call into librqcd.so, i.e. a kernel launch on the GPU; the
// single-object methods are batches of one. There is no CPU implementation behind these classes: without a
// B200 the first call throws rqcd::Error(RQCD_ERR_NO_DEVICE).  The classes hold data only -- evaluating a
// trajectory, tangent planes etc. on the host is deliberately not offered here.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <limits>
#include <memory>
#include <ostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../rqcd.h"

namespace rqcd {

class Error : public std::runtime_error {
 public:
  Error(int status, const std::string& detail)
      : std::runtime_error(std::string("rqcd: ") + rqcd_strerror(status) + (detail.empty() ? "" : " [" + detail + "]")),
        status_(status) {}
  int status() const { return status_; }

 private:
  int status_;
};

// RAII owner of one rqcd_ctx (one GPU, one stream).
class Context {
 public:
  explicit Context(int device = 0) {
    int st = rqcd_create(&ctx_, device);
    if (st != RQCD_OK) throw Error(st, "rqcd_create");
  }
  ~Context() { rqcd_destroy(ctx_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  rqcd_ctx* get() const { return ctx_; }
  void check(int st, bool allow_depth_cap = false) const {
    if (st == RQCD_OK || (allow_depth_cap && st == RQCD_ERR_DEPTH_CAP)) return;
    throw Error(st, rqcd_last_error(ctx_));
  }
  // The context Batch users may share (created on first use, device 0 or $RQCD_DEVICE).
  static Context& Default() {
    static Context c(getenv("RQCD_DEVICE") ? atoi(getenv("RQCD_DEVICE")) : 0);
    return c;
  }
  // The context of the single-object calls (RapidTrajectoryGenerator::Generate, CollisionChecker::CollisionCheck, ...):
  // private to them, because CollisionCheck(obstacle, minT) uploads a one-entry obstacle table for every call and must
  // not replace the table a Batch on the shared context holds.
  static Context& Single() {
    static Context c(getenv("RQCD_DEVICE") ? atoi(getenv("RQCD_DEVICE")) : 0);
    return c;
  }

 private:
  rqcd_ctx* ctx_ = nullptr;
};

}  // namespace rqcd

namespace CommonMath {

struct Vec3 {
  double x, y, z;
  Vec3() : x(std::numeric_limits<double>::quiet_NaN()), y(x), z(x) {}
  Vec3(double xi, double yi, double zi) : x(xi), y(yi), z(zi) {}
  double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
  double& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
  Vec3 operator-(const Vec3& r) const { return Vec3(x - r.x, y - r.y, z - r.z); }  // Vec3.hpp:126-128
  Vec3 operator/(double d) const { return Vec3(x / d, y / d, z / d); }            // Vec3.hpp:164-166
  double GetNorm2Squared() const { return x * x + y * y + z * z; }                 // Vec3.hpp:103-109 (Dot: x*x + y*y + z*z)
  double GetNorm2() const { return std::sqrt(GetNorm2Squared()); }                 // Vec3.hpp:112-114
};

// Unit quaternion (w, x, y, z), local -> global, as Rotation(a, b, c, d) of the reference (Rotation.hpp:42-47).
struct Rotation {
  double v[4];
  Rotation(double a, double b, double c, double d) : v{a, b, c, d} {}
  static Rotation Identity() { return Rotation(1, 0, 0, 0); }
  static Rotation FromAxisAngle(const Vec3& unitAxis, double angle) {  // Rotation.hpp:92-97 (parameter set-up only)
    const double s = std::sin(angle * 0.5);
    return Rotation(std::cos(angle * 0.5), s * unitAxis.x, s * unitAxis.y, s * unitAxis.z);
  }
  // direction = axis, magnitude = angle [rad]; below one arc second: identity (Rotation.hpp:82-87)
  static Rotation FromRotationVector(const Vec3& rotVec) {
    const double theta = rotVec.GetNorm2();
    if (theta < 4.84813681e-6) return Identity();
    return FromAxisAngle(rotVec / theta, theta);
  }
  Rotation Inverse() const { return Rotation(v[0], -v[1], -v[2], -v[3]); }  // Rotation.hpp:63-65
};

struct Boundary {  // ConvexObj.hpp:29-53
  Vec3 point, normal;
};

// The reference's ConvexObj is an open virtual interface (GetTangentPlane / IsPointInside). On the GPU the
// set is closed: a ConvexObj here is anything that can describe itself as an rqcd_obstacle record.
class ConvexObj {
 public:
  virtual ~ConvexObj() {}
  virtual rqcd_obstacle Record() const = 0;
};

class Sphere : public ConvexObj {  // Sphere.hpp:40-72
 public:
  Sphere(Vec3 center, double radius) : c_(center), r_(radius) {}
  rqcd_obstacle Record() const override {
    rqcd_obstacle o = {};
    o.type = RQCD_SPHERE;
    o.center[0] = c_.x, o.center[1] = c_.y, o.center[2] = c_.z;
    o.radius = r_;
    o.quat[0] = 1;
    return o;
  }

 private:
  Vec3 c_;
  double r_;
};

class RectPrism : public ConvexObj {  // RectPrism.hpp:46-123
 public:
  RectPrism(Vec3 center, Vec3 sideLengths, Rotation rotation) : c_(center), s_(sideLengths), q_(rotation) {}
  void SetCenterPoint(Vec3 c) { c_ = c; }
  void SetSideLengths(Vec3 s) { s_ = s; }
  void SetRotation(Rotation r) { q_ = r; }
  rqcd_obstacle Record() const override {
    rqcd_obstacle o = {};
    o.type = RQCD_RECT_PRISM;
    o.center[0] = c_.x, o.center[1] = c_.y, o.center[2] = c_.z;
    o.side[0] = s_.x, o.side[1] = s_.y, o.side[2] = s_.z;
    for (int i = 0; i < 4; i++) o.quat[i] = q_.v[i];
    return o;
  }

 private:
  Vec3 c_, s_;
  Rotation q_;
};

// A non-rotating obstacle translating along its own quintic: the composition
// CollisionChecker(traj - obstacleTraj).CollisionCheck(shape, minT) of Trajectory.hpp:121-128.
class Trajectory;
class MovingObj : public ConvexObj {
 public:
  MovingObj(std::shared_ptr<ConvexObj> shapeAtOrigin, const Trajectory& motion);
  rqcd_obstacle Record() const override { return rec_; }

 private:
  rqcd_obstacle rec_;
};

// Quintic coefficients (index 0 <-> t^5) and the time window; Trajectory.hpp:44-69.
class Trajectory {
 public:
  Trajectory(Vec3 c0, Vec3 c1, Vec3 c2, Vec3 c3, Vec3 c4, Vec3 c5, double startTime, double endTime)
      : c_{c0, c1, c2, c3, c4, c5}, t0_(startTime), t1_(endTime) {}
  Vec3 operator[](int i) const { return c_[i]; }
  double GetStartTime() const { return t0_; }
  double GetEndTime() const { return t1_; }
  // The relative position of two trajectories on their common window (Trajectory.hpp:121-128): coefficient-wise
  // difference, max of the start times, min of the end times. Data arithmetic only -- the check of the result against
  // a shape is CollisionChecker's; for MANY moving obstacles use MovingObj in a Batch, where the device forms the
  // difference per (trajectory, obstacle) pair.
  Trajectory operator-(const Trajectory rhs) const {
    const double startTime = t0_ < rhs.t0_ ? rhs.t0_ : t0_;  // std::max(_startTime, rhs.GetStartTime())
    const double endTime = rhs.t1_ < t1_ ? rhs.t1_ : t1_;    // std::min(_endTime, rhs.GetEndTime())
    return Trajectory(c_[0] - rhs.c_[0], c_[1] - rhs.c_[1], c_[2] - rhs.c_[2], c_[3] - rhs.c_[3], c_[4] - rhs.c_[4],
                      c_[5] - rhs.c_[5], startTime, endTime);
  }

 private:
  Vec3 c_[6];
  double t0_, t1_;
};

inline MovingObj::MovingObj(std::shared_ptr<ConvexObj> shape, const Trajectory& motion) : rec_(shape->Record()) {
  rec_.moving = 1;
  for (int k = 0; k < 6; k++)
    for (int d = 0; d < 3; d++) rec_.motion[3 * k + d] = motion[k][d];
  rec_.motion_t_start = motion.GetStartTime();
  rec_.motion_t_end = motion.GetEndTime();
}

}  // namespace CommonMath

namespace rqcd {

// N candidates x M obstacles: the batched call. Boundary conditions are structure-of-arrays (rqcd_bc_soa).
class Batch {
 public:
  explicit Batch(Context& ctx) : ctx_(ctx) {}

  void SetObstacles(const std::vector<std::shared_ptr<CommonMath::ConvexObj>>& obstacles) {
    std::vector<rqcd_obstacle> rec;
    for (auto& o : obstacles) rec.push_back(o->Record());
    ctx_.check(rqcd_set_obstacles(ctx_.get(), rec.data(), (int32_t)rec.size()));
    m_ = (int32_t)rec.size();
  }

  struct Result {
    std::vector<uint8_t> verdict;         // CollisionChecker::CollisionResult per candidate (first hit in index order)
    std::vector<int32_t> first_obstacle;  // -1 when collision free
    std::vector<double> cost;             // RapidTrajectoryGenerator::GetCost()
    std::vector<double> coeffs;           // only when asked for: RQCD_COEFF_ROWS x n, GetTrajectory() coefficients
    rqcd_summary summary;
  };

  // bc: RQCD_BC_ROWS x n, row r at bc + r*n (host memory). earlyExit = ForestPerformanceEval's inner break.
  Result GenerateAndCheck(const double* bc, int64_t n, double minTimeSection, uint32_t goalMask = RQCD_GOAL_ALL,
                          bool earlyExit = false, int64_t indexBase = 0, bool wantCoeffs = false) {
    Result r;
    r.verdict.resize(n);
    r.first_obstacle.resize(n);
    r.cost.resize(n);
    rqcd_bc_soa in = {bc, n, goalMask, 0};
    rqcd_out out = {};
    out.verdict = r.verdict.data();
    out.first_obstacle = r.first_obstacle.data();
    out.cost = r.cost.data();
    out.summary = &r.summary;
    if (wantCoeffs) {
      r.coeffs.resize((size_t)RQCD_COEFF_ROWS * n);
      out.coeffs = r.coeffs.data();
      out.coeff_stride = n;
    }
    ctx_.check(rqcd_generate_check(ctx_.get(), &in, n, minTimeSection, earlyExit ? RQCD_F_EARLY_EXIT : 0, indexBase, &out),
               true);
    return r;
  }

  // PerformanceEval shape: candidate i against its own sphere i (spheres: 4 rows cx, cy, cz, r).
  Result GenerateAndCheckPairwise(const double* bc, const double* spheres, int64_t n, double minTimeSection,
                                  uint32_t goalMask = RQCD_GOAL_ALL) {
    Result r;
    r.verdict.resize(n);
    r.cost.resize(n);
    rqcd_bc_soa in = {bc, n, goalMask, 0};
    rqcd_out out = {};
    out.verdict = r.verdict.data();
    out.cost = r.cost.data();
    out.summary = &r.summary;
    ctx_.check(rqcd_generate_check_pairwise(ctx_.get(), &in, spheres, n, n, minTimeSection, 0, 0, &out), true);
    return r;
  }

  // RapidTrajectoryGenerator::CheckInputFeasibility for every candidate. group (n entries or nullptr): candidates with
  // equal consecutive group ids share one generator object, as ForestPerformanceEval's batches do.
  std::vector<uint8_t> InputFeasibility(const double* bc, int64_t n, const CommonMath::Vec3& gravity, double fminAllowed,
                                        double fmaxAllowed, double wmaxAllowed, double minTimeSection,
                                        const int64_t* group = nullptr, uint32_t goalMask = RQCD_GOAL_ALL) {
    std::vector<uint8_t> res(n);
    rqcd_bc_soa in = {bc, n, goalMask, 0};
    const double g[3] = {gravity.x, gravity.y, gravity.z};
    ctx_.check(rqcd_check_input_feasibility(ctx_.get(), &in, n, g, fminAllowed, fmaxAllowed, wmaxAllowed, minTimeSection,
                                            group, 0, res.data()));
    return res;
  }

  // The planner's candidate loop (rqcd_plan_best): drop candidates that cost more than maxCost, that are not input
  // feasible, that cross one of `planes` or that hit an obstacle; summary.best_* is the cheapest survivor.
  struct Plan {
    std::vector<uint8_t> stage;  // rqcd_plan_stage per candidate
    rqcd_summary summary;
  };
  Plan PlanBest(const double* bc, int64_t n, double minTimeSection, double maxCost, const CommonMath::Vec3& gravity,
                double fminAllowed, double fmaxAllowed, double wmaxAllowed,
                const std::vector<CommonMath::Boundary>& planes = {}, uint32_t goalMask = RQCD_GOAL_ALL) {
    Plan r;
    r.stage.resize(n);
    std::vector<double> pl;
    for (auto& b : planes)
      for (double v : {b.point.x, b.point.y, b.point.z, b.normal.x, b.normal.y, b.normal.z}) pl.push_back(v);
    rqcd_plan_params prm = {};
    prm.max_cost = maxCost;
    prm.check_input = 1;
    prm.gravity[0] = gravity.x, prm.gravity[1] = gravity.y, prm.gravity[2] = gravity.z;
    prm.fmin = fminAllowed, prm.fmax = fmaxAllowed, prm.wmax = wmaxAllowed;
    prm.planes = pl.empty() ? nullptr : pl.data();
    prm.n_planes = (int32_t)planes.size();
    prm.min_time_section = minTimeSection;
    rqcd_bc_soa in = {bc, n, goalMask, 0};
    ctx_.check(rqcd_plan_best(ctx_.get(), &in, n, &prm, 0, 0, r.stage.data(), &r.summary), true);
    return r;
  }

  int32_t NumObstacles() const { return m_; }

 private:
  Context& ctx_;
  int32_t m_ = 0;
};

// Several GPUs behind ONE C++ host (rqcd_group_*): candidates sharded by index over the devices, the obstacle set on every
// device, each shard fed by its own host thread through the streaming pipeline; the shards' verdict counts and best
// candidates are merged on the first device (peer-mapped slot table over NVLink + a merge kernel) -- the reduction the
// drivers do on the host after every check (ForestPerformanceEval.cpp:216-230). Same results for every device count.
class Group {
 public:
  explicit Group(const std::vector<int32_t>& devices) {
    const int st = rqcd_group_create(&g_, devices.data(), (int32_t)devices.size());
    if (st != RQCD_OK) throw Error(st, "rqcd_group_create");
  }
  ~Group() {
    if (g_) rqcd_group_destroy(g_);
  }
  Group(const Group&) = delete;
  Group& operator=(const Group&) = delete;
  int32_t Size() const { return rqcd_group_size(g_); }
  bool PeerSlots() const { return rqcd_group_peer_slots(g_) != 0; }

  void SetObstacles(const std::vector<std::shared_ptr<CommonMath::ConvexObj>>& obstacles) {
    std::vector<rqcd_obstacle> rec;
    for (auto& o : obstacles) rec.push_back(o->Record());
    check(rqcd_group_set_obstacles(g_, rec.data(), (int32_t)rec.size()));
  }

  // Batch::GenerateAndCheck over all devices: host buffers in, host buffers out, merged summary.
  Batch::Result GenerateAndCheck(const double* bc, int64_t n, double minTimeSection, uint32_t goalMask = RQCD_GOAL_ALL,
                                 bool earlyExit = false) {
    Batch::Result r;
    r.verdict.resize(n);
    r.first_obstacle.resize(n);
    r.cost.resize(n);
    rqcd_bc_soa in = {bc, n, goalMask, 0};
    rqcd_out out = {};
    out.verdict = r.verdict.data();
    out.first_obstacle = r.first_obstacle.data();
    out.cost = r.cost.data();
    out.summary = &r.summary;
    check(rqcd_group_generate_check(g_, &in, n, minTimeSection, earlyExit ? RQCD_F_EARLY_EXIT : 0, &out), true);
    return r;
  }

 private:
  void check(int st, bool allow_depth_cap = false) const {
    if (st == RQCD_OK || (allow_depth_cap && st == RQCD_ERR_DEPTH_CAP)) return;
    throw Error(st, rqcd_group_last_error(g_));
  }
  rqcd_group* g_ = nullptr;
};

}  // namespace rqcd

namespace rqcd {

// The on-disk formats of the reference's Forest driver, so that scripts/plotSingleForestBatch.py reads GPU output:
// one row per candidate = 18 coefficients (t^5..t^0, xyz each), Tf, inputFeas, stateFeas, three per-call timings
// [ns], every field followed by a comma, default ostream formatting (test/ForestPerformanceEval.cpp:255-267).
// timings may be nullptr (a batched GPU call has no per-candidate times): zeros are written.
inline void WriteForestBatchCsv(std::ostream& os, int64_t n, const double* coeffs, int64_t coeffStride, const double* Tf,
                                const uint8_t* inputFeas, const uint8_t* stateFeas, const long* timingsNs = nullptr) {
  for (int64_t i = 0; i < n; i++) {
    for (int k = 0; k < RQCD_COEFF_ROWS; k++) os << coeffs[k * coeffStride + i] << ",";
    os << Tf[i] << "," << (int)inputFeas[i] << "," << (int)stateFeas[i] << ",";
    for (int k = 0; k < 3; k++) os << (timingsNs ? timingsNs[3 * i + k] : 0L) << ",";
    os << "\n";
  }
}

// One row per obstacle: centre, side lengths, row-major rotation matrix (ForestPerformanceEval.cpp:139-147).
inline void WriteTreeParamsCsv(std::ostream& os, const std::vector<std::shared_ptr<CommonMath::ConvexObj>>& trees) {
  for (auto& t : trees) {
    const rqcd_obstacle o = t->Record();
    const double* v = o.quat;  // printed to 6 digits: the matrix of Rotation.hpp:203-220
    const double r0 = v[0] * v[0], r1 = v[1] * v[1], r2 = v[2] * v[2], r3 = v[3] * v[3];
    const double R[9] = {r0 + r1 - r2 - r3,
                         2 * v[1] * v[2] - 2 * v[0] * v[3],
                         2 * v[1] * v[3] + 2 * v[0] * v[2],
                         2 * v[1] * v[2] + 2 * v[0] * v[3],
                         r0 - r1 + r2 - r3,
                         2 * v[2] * v[3] - 2 * v[0] * v[1],
                         2 * v[1] * v[3] - 2 * v[0] * v[2],
                         2 * v[2] * v[3] + 2 * v[0] * v[1],
                         r0 - r1 - r2 + r3};
    for (int i = 0; i < 3; i++) os << o.center[i] << ",";
    for (int i = 0; i < 3; i++) os << o.side[i] << ",";
    for (int i = 0; i < 9; i++) os << R[i] << ",";
    os << "\n";
  }
}

}  // namespace rqcd

namespace RapidQuadrocopterTrajectoryGenerator {

class RapidTrajectoryGenerator {
 public:
  enum InputFeasibilityResult {  // RapidTrajectoryGenerator.hpp:63-69 = rqcd_input_feasibility
    InputFeasible = 0,
    InputIndeterminable = 1,
    InputInfeasibleThrustHigh = 2,
    InputInfeasibleThrustLow = 3,
    InputInfeasibleRates = 4,
  };
  enum StateFeasibilityResult { StateFeasible = 0, StateInfeasible = 1 };  // :71-74

  RapidTrajectoryGenerator(const CommonMath::Vec3 x0, const CommonMath::Vec3 v0, const CommonMath::Vec3 a0,
                           const CommonMath::Vec3 gravity)
      : grav_(gravity) {
    for (int i = 0; i < 3; i++) bc_[RQCD_BC_P0 + i] = x0[i], bc_[RQCD_BC_V0 + i] = v0[i], bc_[RQCD_BC_A0 + i] = a0[i];
    Reset();
  }
  void SetGoalPosition(const CommonMath::Vec3 in) { for (unsigned i = 0; i < 3; i++) SetGoalPositionInAxis(i, in[i]); }
  void SetGoalVelocity(const CommonMath::Vec3 in) { for (unsigned i = 0; i < 3; i++) SetGoalVelocityInAxis(i, in[i]); }
  void SetGoalAcceleration(const CommonMath::Vec3 in) { for (unsigned i = 0; i < 3; i++) SetGoalAccelerationInAxis(i, in[i]); }
  void SetGoalPositionInAxis(const unsigned ax, const double in) { mask_ |= RQCD_GOAL_POS(ax), bc_[RQCD_BC_PF + ax] = in; }
  void SetGoalVelocityInAxis(const unsigned ax, const double in) { mask_ |= RQCD_GOAL_VEL(ax), bc_[RQCD_BC_VF + ax] = in; }
  void SetGoalAccelerationInAxis(const unsigned ax, const double in) { mask_ |= RQCD_GOAL_ACC(ax), bc_[RQCD_BC_AF + ax] = in; }
  //! Clears the end-state constraints (goal flags persist across Generate calls, as in the reference) and the
  //! acceleration peak-time cache of CheckInputFeasibility (SingleAxisTrajectory::Reset, SingleAxisTrajectory.cpp:41).
  void Reset() {
    mask_ = 0;
    for (int i = RQCD_BC_PF; i < RQCD_BC_TF; i++) bc_[i] = 0;
    bc_[RQCD_BC_TF] = 0;
    cost_ = std::numeric_limits<double>::quiet_NaN();
    for (double& c : coeffs_) c = std::numeric_limits<double>::quiet_NaN();
    for (double& c : abg_) c = std::numeric_limits<double>::quiet_NaN();
    feasHistory_.clear();
  }
  //! One rqcd_generate launch with n = 1.
  void Generate(const double timeToGo) {
    bc_[RQCD_BC_TF] = timeToGo;
    rqcd_bc_soa in = {bc_, 1, mask_, 0};
    rqcd_out out = {};
    out.cost = &cost_;
    out.coeffs = coeffs_;
    out.coeff_stride = 1;
    out.params = abg_;
    out.param_stride = 1;
    rqcd::Context& c = rqcd::Context::Single();
    c.check(rqcd_generate(c.get(), &in, 1, 0, &out));
  }
  double GetCost() const { return cost_; }
  double GetFinalTime() const { return bc_[RQCD_BC_TF]; }
  //! The parameters the generator computed on the device (rqcd_out::params), not back-computed from the coefficients.
  double GetAxisParamAlpha(int i) const { return abg_[0 + i]; }
  double GetAxisParamBeta(int i) const { return abg_[3 + i]; }
  double GetAxisParamGamma(int i) const { return abg_[6 + i]; }
  CommonMath::Trajectory GetTrajectory() const {
    using CommonMath::Vec3;
    const double* k = coeffs_;
    return CommonMath::Trajectory(Vec3(k[0], k[1], k[2]), Vec3(k[3], k[4], k[5]), Vec3(k[6], k[7], k[8]),
                                  Vec3(k[9], k[10], k[11]), Vec3(k[12], k[13], k[14]), Vec3(k[15], k[16], k[17]), 0,
                                  bc_[RQCD_BC_TF]);
  }

  //! rqcd_check_input_feasibility for this object (RapidTrajectoryGenerator.cpp:74-160). The reference caches each
  //! axis' acceleration peak times in the object from the first call that needed them until Reset()
  //! (SingleAxisTrajectory.cpp:133-158), so a loop that re-Generates on ONE object evaluates later candidates with the
  //! first one's peak times (ForestPerformanceEval.cpp:177-209). Reproduced here through the library's `group`
  //! argument: the object remembers the boundary conditions of every call since Reset() and submits them as one group;
  //! the answer is the last member's. (A loop over many candidates is what rqcd::Batch::InputFeasibility is for.)
  InputFeasibilityResult CheckInputFeasibility(double fminAllowed, double fmaxAllowed, double wmaxAllowed,
                                               double minTimeSection) {
    feasHistory_.insert(feasHistory_.end(), bc_, bc_ + RQCD_BC_ROWS);
    const int64_t k = (int64_t)(feasHistory_.size() / RQCD_BC_ROWS);
    std::vector<double> soa((size_t)RQCD_BC_ROWS * k);
    for (int64_t j = 0; j < k; j++)
      for (int r = 0; r < RQCD_BC_ROWS; r++) soa[(size_t)r * k + j] = feasHistory_[(size_t)j * RQCD_BC_ROWS + r];
    std::vector<int64_t> group(k, 0);
    std::vector<uint8_t> res(k);
    rqcd_bc_soa in = {soa.data(), k, mask_, 0};
    const double g[3] = {grav_.x, grav_.y, grav_.z};
    rqcd::Context& c = rqcd::Context::Single();
    c.check(rqcd_check_input_feasibility(c.get(), &in, k, g, fminAllowed, fmaxAllowed, wmaxAllowed, minTimeSection,
                                         group.data(), 0, res.data()));
    return (InputFeasibilityResult)res[k - 1];
  }

  //! rqcd_generate_check_planes with one plane (RapidTrajectoryGenerator.cpp:162-212, the same test as
  //! CollisionChecker::CollisionCheck(Boundary)): StateInfeasible when the trajectory reaches the plane's wrong side.
  StateFeasibilityResult CheckPositionFeasibility(CommonMath::Vec3 boundaryPoint, CommonMath::Vec3 boundaryNormal) {
    const double plane[6] = {boundaryPoint.x, boundaryPoint.y, boundaryPoint.z, boundaryNormal.x, boundaryNormal.y, boundaryNormal.z};
    rqcd_bc_soa in = {bc_, 1, mask_, 0};
    uint8_t hit = 0;
    rqcd::Context& c = rqcd::Context::Single();
    c.check(rqcd_generate_check_planes(c.get(), &in, 1, plane, 1, 0, &hit, nullptr));
    return hit ? StateInfeasible : StateFeasible;
  }

  static const char* GetInputFeasibilityResultName(InputFeasibilityResult fr) {  // :243-257
    switch (fr) {
      case InputFeasible: return "Feasible";
      case InputIndeterminable: return "Indeterminable";
      case InputInfeasibleThrustHigh: return "InfeasibleThrustHigh";
      case InputInfeasibleThrustLow: return "InfeasibleThrustLow";
      case InputInfeasibleRates: return "InfeasibleRates";
    }
    return "Unknown!";
  }

 private:
  double bc_[RQCD_BC_ROWS] = {0};
  uint32_t mask_ = 0;
  double cost_;
  double coeffs_[RQCD_COEFF_ROWS];
  double abg_[RQCD_PARAM_ROWS];
  std::vector<double> feasHistory_;  // RQCD_BC_ROWS doubles per CheckInputFeasibility call since Reset()
  CommonMath::Vec3 grav_;
};

}  // namespace RapidQuadrocopterTrajectoryGenerator

namespace RapidCollisionChecker {

class CollisionChecker {
 public:
  enum CollisionResult { NoCollision = 0, Collision = 1, CollisionIndeterminable = 2 };

  explicit CollisionChecker(CommonMath::Trajectory traj) : traj_(traj) {}

  //! One rqcd_check_planes launch with n = 1, p = 1 (CollisionChecker.hpp:49, src/CollisionChecker.cpp:25-68; see rqcd.h
  //! for what "crosses the plane" means where the reference's own code reads uninitialised roots).
  CollisionResult CollisionCheck(CommonMath::Boundary boundary) {
    const double plane[6] = {boundary.point.x,  boundary.point.y,  boundary.point.z,
                             boundary.normal.x, boundary.normal.y, boundary.normal.z};
    double k[RQCD_COEFF_ROWS];
    for (int i = 0; i < 6; i++)
      for (int d = 0; d < 3; d++) k[3 * i + d] = traj_[i][d];
    const double t0 = traj_.GetStartTime(), t1 = traj_.GetEndTime();
    rqcd_coeff_soa in = {k, 1, &t0, &t1};
    uint8_t hit = 0;
    rqcd::Context& c = rqcd::Context::Single();
    c.check(rqcd_check_planes(c.get(), &in, 1, plane, 1, 0, &hit, nullptr));
    return hit ? Collision : NoCollision;
  }

  //! One rqcd_set_obstacles + rqcd_check launch with n = 1, m = 1 on the single-object context (never the context a
  //! Batch holds its obstacle table on).
  CollisionResult CollisionCheck(std::shared_ptr<CommonMath::ConvexObj> obstacle, double minTimeSection) {
    rqcd::Context& c = rqcd::Context::Single();
    rqcd_obstacle rec = obstacle->Record();
    c.check(rqcd_set_obstacles(c.get(), &rec, 1));
    double k[RQCD_COEFF_ROWS];
    for (int i = 0; i < 6; i++)
      for (int d = 0; d < 3; d++) k[3 * i + d] = traj_[i][d];
    const double t0 = traj_.GetStartTime(), t1 = traj_.GetEndTime();
    rqcd_coeff_soa in = {k, 1, &t0, &t1};
    uint8_t verdict = 0;
    rqcd_out out = {};
    out.verdict = &verdict;
    c.check(rqcd_check(c.get(), &in, nullptr, 1, minTimeSection, 0, 0, &out));
    return (CollisionResult)verdict;
  }

  static const char* GetCollisionResultName(CollisionResult fr) {
    switch (fr) {
      case NoCollision: return "No Collision";
      case Collision: return "Collision";
      case CollisionIndeterminable: return "Collision Indeterminable";
    }
    return "Unknown!";
  }

 private:
  CommonMath::Trajectory traj_;
};

}  // namespace RapidCollisionChecker
