/*
 * rqcd.h -- C ABI of the B200-native batched motion-primitive generator and
 * convex-obstacle collision checker.
 *
 * This is the drop-in boundary for ONE path of
 * nlbucki/RapidQuadcopterCollisionDetection: the per-trial body of its two
 * drivers (test/PerformanceEval.cpp:140-182, test/ForestPerformanceEval.cpp:177-225),
 * i.e.  RapidTrajectoryGenerator::Generate -> GetTrajectory ->
 * CollisionChecker::CollisionCheck(obstacle, minTimeSection).
 * The reference has no FFI of its own (it is a C++ class library,
 * src/CMakeLists.txt:1-3); each entry point below names the reference
 * interface it replaces.  The C++ facade in include/rqcd/ keeps the reference's
 * class names and calls only these functions.
 *
 * Conventions
 *  - plain pointers and sizes only; the caller owns every buffer.
 *  - every function returns an rqcd_status; nothing throws, nothing aborts.
 *  - all arithmetic is IEEE-754 binary64, evaluated in the reference's
 *    operation order without FMA contraction ("exact" mode).
 *  - there is no CPU fallback: without a CUDA device rqcd_create fails with
 *    RQCD_ERR_NO_DEVICE and nothing else can be called.
 *  - a context is bound to one GPU and one stream; use one context per host
 *    thread / per rank (one process per GPU under torchrun), or an rqcd_group
 *    of contexts for several GPUs in one process.
 */
#ifndef RQCD_H_
#define RQCD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RQCD_VERSION_MAJOR 0
#define RQCD_VERSION_MINOR 1

typedef enum rqcd_status {
  RQCD_OK = 0,
  RQCD_ERR_INVALID_ARG = 1, /* null pointer, n < 0, stride < n, minT <= 0 or NaN, bad obstacle */
  RQCD_ERR_CUDA = 2,        /* a CUDA runtime call failed; see rqcd_last_error */
  RQCD_ERR_NO_DEVICE = 3,   /* no usable sm_100 device: there is no CPU fallback */
  RQCD_ERR_DEPTH_CAP = 4,   /* some pair hit the internal interval-stack cap (verdict 3) */
  RQCD_ERR_NOMEM = 5,
  RQCD_ERR_UNSUPPORTED = 6  /* e.g. obstacle type outside the closed set {sphere, prism} */
} rqcd_status;

/* Verdict of one check. Values 0..2 are the reference's
 * CollisionChecker::CollisionResult (include/RapidCollisionDetection/CollisionChecker.hpp:33-37).
 * 3 never occurs unless (t_end - t_start)/minT exceeds 2^RQCD_MAX_DEPTH. */
typedef enum rqcd_verdict {
  RQCD_NO_COLLISION = 0,
  RQCD_COLLISION = 1,
  RQCD_INDETERMINABLE = 2,
  RQCD_DEPTH_CAP = 3
} rqcd_verdict;

/* RapidTrajectoryGenerator::InputFeasibilityResult
 * (include/RapidQuadcopterTrajectories/RapidTrajectoryGenerator.hpp:63-69). */
typedef enum rqcd_input_feasibility {
  RQCD_INPUT_FEASIBLE = 0,
  RQCD_INPUT_INDETERMINABLE = 1,
  RQCD_INPUT_INFEASIBLE_THRUST_HIGH = 2,
  RQCD_INPUT_INFEASIBLE_THRUST_LOW = 3,
  RQCD_INPUT_INFEASIBLE_RATES = 4,
  RQCD_INPUT_DEPTH_CAP = 5 /* not a reference value: more than RQCD_MAX_DEPTH pending second halves (needs Tf / minTimeSection
                              > 2^40); the reference would recurse on. Never silent: reported instead of folded into 1 */
} rqcd_input_feasibility;

/* ---- obstacles -----------------------------------------------------------
 * The reference's open virtual interface CommonMath::ConvexObj
 * (include/CommonMath/ConvexObj.hpp:61-82) becomes a closed tagged record:
 * Sphere (include/CommonMath/Sphere.hpp:40-72) or RectPrism
 * (include/CommonMath/RectPrism.hpp:46-100). */
typedef enum rqcd_obstacle_type { RQCD_SPHERE = 0, RQCD_RECT_PRISM = 1 } rqcd_obstacle_type;

typedef struct rqcd_obstacle {
  int32_t type;      /* rqcd_obstacle_type */
  int32_t moving;    /* 0: static.  1: translates (no rotation) along `motion`; the pair check
                        is CollisionChecker(traj - motion).CollisionCheck(shape, minT), the
                        composition of include/CommonMath/Trajectory.hpp:121-128 */
  double center[3];  /* Sphere::_centerPoint / RectPrism::_centerPoint (for a moving obstacle:
                        the shape's centre in the frame that moves with `motion`) */
  double radius;     /* sphere only, > 0 (Sphere.hpp:44) */
  double side[3];    /* prism only: full side lengths, each > 0 (RectPrism.hpp:52) */
  double quat[4];    /* prism only: Rotation(a,b,c,d) local->global (Rotation.hpp:42-47) */
  double motion[18]; /* moving only: quintic coefficients, motion[3*k+dim], k=0 is t^5
                        (Trajectory.hpp:26-31) */
  double motion_t_start, motion_t_end; /* moving only: window of `motion` */
} rqcd_obstacle;

/* ---- trajectories --------------------------------------------------------
 * Boundary conditions, structure-of-arrays: row r of the matrix starts at
 * data + r*stride and holds n doubles.  Rows follow the argument order of
 * RapidTrajectoryGenerator(x0,v0,a0,gravity) + SetGoalPosition/Velocity/
 * Acceleration + Generate(Tf) (RapidTrajectoryGenerator.hpp:77-114). */
enum {
  RQCD_BC_P0 = 0,  /* rows 0..2  : initial position x,y,z */
  RQCD_BC_V0 = 3,  /* rows 3..5  : initial velocity */
  RQCD_BC_A0 = 6,  /* rows 6..8  : initial acceleration */
  RQCD_BC_PF = 9,  /* rows 9..11 : goal position   (ignored where the goal bit is clear) */
  RQCD_BC_VF = 12, /* rows 12..14: goal velocity */
  RQCD_BC_AF = 15, /* rows 15..17: goal acceleration */
  RQCD_BC_TF = 18, /* row 18     : duration Tf */
  RQCD_BC_ROWS = 19
};

/* goal_mask bit (3*axis + c): component c (0 pos, 1 vel, 2 acc) of axis is constrained
 * (SingleAxisTrajectory::SetGoalPosition/Velocity/Acceleration, SingleAxisTrajectory.hpp:39-45).
 * One mask per batch, as in both reference drivers. */
#define RQCD_GOAL_POS(axis) (1u << (3 * (axis) + 0))
#define RQCD_GOAL_VEL(axis) (1u << (3 * (axis) + 1))
#define RQCD_GOAL_ACC(axis) (1u << (3 * (axis) + 2))
#define RQCD_GOAL_ALL 0x1FFu

typedef struct rqcd_bc_soa {
  const double* data; /* RQCD_BC_ROWS rows */
  int64_t stride;     /* doubles between rows, >= first + n */
  uint32_t goal_mask;
  /* Rows that candidates share. The reference's Forest loop draws ONE initial state per batch of 100 candidates and
   * varies only the goal position and the duration (test/ForestPerformanceEval.cpp:175-193); such a batch need not
   * carry 19 doubles per candidate. Bit r set: row r holds one value per GROUP of group_size consecutive candidates
   * (candidate i reads element i / group_size); with host pointers only those elements cross PCIe. 0 (default):
   * every row is per candidate. */
  uint32_t shared_rows;
  int64_t group_size; /* >= 1 when shared_rows != 0 (a value >= first + n makes a shared row one constant) */
  int64_t first;      /* the call's candidate 0 is element `first` of the per-candidate rows and belongs to group
                         first / group_size: lets shards of one batch (rqcd_group_*, ranks) share one matrix. Default 0 */
} rqcd_bc_soa;

/* Pre-built quintics (CommonMath::Trajectory, Trajectory.hpp:44-69), structure-of-arrays:
 * row 3*k+dim (k = 0 is t^5) starts at coeffs + (3*k+dim)*stride. */
#define RQCD_COEFF_ROWS 18
typedef struct rqcd_coeff_soa {
  const double* coeffs;
  int64_t stride;
  const double* t_start; /* n doubles, or NULL for 0 */
  const double* t_end;   /* n doubles, required */
} rqcd_coeff_soa;

/* ---- results ------------------------------------------------------------- */
typedef struct rqcd_summary {
  uint64_t traj_counts[4]; /* trajectories by first-hit verdict (index = rqcd_verdict) */
  uint64_t pair_counts[4]; /* executed trajectory-obstacle checks by verdict */
  uint64_t pairs_checked;  /* = sum(pair_counts); n*m unless RQCD_F_EARLY_EXIT */
  double best_cost;        /* min GetCost() over NoCollision trajectories, +inf if none */
  int64_t best_index;      /* index_base + local index of that trajectory (lowest on ties), -1 if none */
} rqcd_summary;

typedef struct rqcd_out {
  uint8_t* verdict;        /* [n] first obstacle in index order whose check != NoCollision gives
                              the verdict, else NoCollision (ForestPerformanceEval.cpp:216-225); may be NULL */
  int32_t* first_obstacle; /* [n] index of that obstacle, -1 if none; may be NULL */
  uint8_t* verdict_matrix; /* [n*m] row-major per trajectory, every pair's own verdict; may be NULL.
                              Not allowed together with RQCD_F_EARLY_EXIT */
  double* cost;            /* [n] RapidTrajectoryGenerator::GetCost(); may be NULL */
  double* coeffs;          /* RQCD_COEFF_ROWS rows, GetTrajectory() coefficients; may be NULL */
  int64_t coeff_stride;
  rqcd_summary* summary;   /* HOST pointer (also with RQCD_F_DEVICE_PTRS); may be NULL. When non-NULL
                              the call synchronises the context's stream before returning. */
  uint8_t* fragile;        /* [n], with RQCD_F_FRAGILE (required then): 1 when the trajectory's verdict or first obstacle
                              may depend on the last bits of the cubic solver's acos / cos / pow, see the flag */
  uint8_t* input_feasibility; /* [n], optional, with rqcd_set_input_filter: each candidate's rqcd_input_feasibility */
  double* params;          /* rqcd_generate only: RQCD_PARAM_ROWS rows alpha x,y,z, beta x,y,z, gamma x,y,z =
                              GetAxisParamAlpha/Beta/Gamma(axis) (RapidTrajectoryGenerator.hpp:219-229); may be NULL */
  int64_t param_stride;
} rqcd_out;
#define RQCD_PARAM_ROWS 9

/* flags */
#define RQCD_F_DEVICE_PTRS 0x1u /* every data pointer in the in/out structs is a device pointer on the
                                   context's GPU; the call only enqueues work on the context's stream */
/* Without RQCD_F_DEVICE_PTRS the data pointers are HOST pointers and the call returns when the results are in the
 * caller's buffers. Page-locked (cudaHostAlloc / cudaHostRegister) input streams in chunk by chunk beside ONE
 * persistent launch (the copy engine advances a progress word the kernel waits on); pageable input is copied
 * completely before the launch. */
#define RQCD_F_EARLY_EXIT 0x2u  /* stop each trajectory at its first non-NoCollision obstacle (what the
                                   Forest driver does); default is all pairs */

/* Diagnostic: mark the results that are NOT pinned by the reference's arithmetic. Every operation of the path is the
 * reference's own IEEE sequence except the transcendental core of Quartic::solveP3 (src/quartic.cpp:50-55, :60: acos, cos,
 * pow), where the reference calls its C library and this library evaluates polynomials (<= 1.5 ulp); glibc itself gives
 * different last bits on different CPUs. Where the closed-form roots are ill-conditioned (grazing contact, coefficient
 * ratios of 1e3 and more) that ulp decides on which side of the tangent plane a computed critical point lands. With this
 * flag a second kernel (k_fragile) walks every pair's recursion again, in lockstep with eight copies whose transcendental
 * outputs are moved by +-2^-50 (the reference's three independent cosines in all sign patterns, Cardano's cube root
 * relatively), and asks of every SIGN decision of every section that can change the outcome -- window vs minTimeSection,
 * midpoint inside, the solver's branches, each root against ts / mid / tf, the plane side of each test point -- whether it
 * is farther from zero than 4 x what the perturbation did to it + 8 ulps of the terms it is computed from (its own
 * rounding noise; a plane-side value on its noise floor is the signature of grazing contact). Decisions that cannot
 * involve the transcendentals (the first section's window / inside / end-point tests) are exempt, and so are positions
 * and existence of test points that provably decide nothing (a free root next to a free end; the double root every
 * rest-to-rest primitive has at t = Tf). out->fragile[i] = 1 when a pair of trajectory i at or before its first hit has an
 * unpinned decision; with a verdict_matrix, bit 7 of an entry marks the pair (the verdict stays in bits 0..1). Verdicts
 * are never changed. An unmarked verdict is the reference's on every conforming libm (tests: every difference from the
 * CPU reference on adversarial grazing populations is marked, with zero tolerance); a marked one MAY be a coin toss in
 * the reference too -- the mark is conservative: 0 of 4e6 Monte Carlo pairs, but 4 % of rest-to-rest primitives, whose
 * double root makes the solver's choice of resolvent root a rounding matter although the test points come out the
 * same. Cost: a diagnostic, not a hot path -- one thread per pair, nine copies, the compiler's divisions: ~6e7 pairs/s
 * (the check itself does 2e10). Tunables, read at rqcd_create: RQCD_FRAGILE_LOG2 (magnitudes, default -50),
 * RQCD_FRAGILE_MARGIN (4), RQCD_FRAGILE_FLOOR (8 ulps). rqcd_fragile_reasons says which decision it was. */
#define RQCD_F_FRAGILE 0x4u

/* Opt-in: run the check with the kernels' FMA-contracted build (the same sources compiled with -fmad=true instead of the
 * default -fmad=false). About 7 % faster; NOT the reference's arithmetic: a contracted a*b+c rounds once where the
 * reference's x86-64 build rounds twice, so coefficients and costs move in their last bits and a verdict can differ
 * from the reference's (measured: 0 of 1.1e9 Monte Carlo pairs, tools/cmp_fmad.py; tests report the count). The
 * input filter, RQCD_F_FRAGILE and every other entry point stay exact. */
#define RQCD_F_FAST_FMA 0x8u

#define RQCD_MAX_DEPTH 40       /* interval-stack entries per check */

typedef struct rqcd_ctx rqcd_ctx;

/* ---- context ------------------------------------------------------------- */
int rqcd_create(rqcd_ctx** ctx, int device);
void rqcd_destroy(rqcd_ctx* ctx);
const char* rqcd_strerror(int status);
const char* rqcd_last_error(const rqcd_ctx* ctx); /* text of the last CUDA failure on this context */
int rqcd_device_info(const rqcd_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, char* name, size_t name_len);

/* The stream the context launches on (a cudaStream_t). rqcd_set_stream lets the caller supply its
 * own (e.g. to time with events); pass NULL to go back to the context-owned stream. */
void* rqcd_get_stream(const rqcd_ctx* ctx);
int rqcd_set_stream(rqcd_ctx* ctx, void* cuda_stream);
int rqcd_synchronize(rqcd_ctx* ctx);
/* Summary of the most recent check launch on this context (synchronises). */
int rqcd_read_summary(rqcd_ctx* ctx, rqcd_summary* out);
/* Number of kernels this context has launched since creation. */
uint64_t rqcd_launch_count(const rqcd_ctx* ctx);
/* Name of the check kernel of the most recent rqcd_generate_check / rqcd_check / rqcd_generate_check_pairwise /
 * rqcd_plan_best launch on this context ("" before the first): what a profile of the call must show. */
const char* rqcd_last_kernel_name(const rqcd_ctx* ctx);
/* RQCD_F_FRAGILE diagnostics: how many pairs this context has marked so far, by the FIRST decision of the pair's recursion
 * that was needed and not pinned: 0 window vs minTimeSection, 1 midpoint inside, 2 quartic or cubic, 3..10 the solver's
 * branches (3 trig / Cardano -- never needed --, 4 |x2| < eps, 5 / 6 choice of the resolvent root, 7 / 8 |D| < eps, 9 / 10 the
 * two quadratic factors' discriminants), 11 + 4 j + {0, 1, 2, 3}: root j (by origin: 0, 1 the first factor's, 2, 3 the
 * second's; the cubic's in 0..2) against ts / mid / tf and its plane side, 27 / 28 plane side of ts / tf, 31 the perturbed
 * copies disagreed on the section's result, its children or the existence of roots. */
int rqcd_fragile_reasons(rqcd_ctx* ctx, uint64_t counts[32]);

/* Replaces constructing Sphere / RectPrism objects and the vector<shared_ptr<ConvexObj>> the drivers
 * hold (ForestPerformanceEval.cpp:130-150). Host pointer always. Rotation matrices and half sides are
 * hoisted here with the reference's expression order (Rotation.hpp:203-220). The set is replicated per
 * GPU; static and moving obstacles may be mixed. */
int rqcd_set_obstacles(rqcd_ctx* ctx, const rqcd_obstacle* obstacles, int32_t m);
int32_t rqcd_num_obstacles(const rqcd_ctx* ctx);

/* ---- the hot path --------------------------------------------------------- */

/* Generate + GetCost + GetTrajectory only (RapidTrajectoryGenerator.cpp:67-72, .hpp:214-241).
 * Uses out->cost, out->coeffs, out->params; other fields ignored. */
int rqcd_generate(rqcd_ctx* ctx, const rqcd_bc_soa* in, int64_t n, uint32_t flags, rqcd_out* out);

/* Fused Generate -> GetTrajectory -> CollisionCheck against the whole obstacle set: the body of the
 * Forest driver's candidate loop (ForestPerformanceEval.cpp:186-225) for n candidates at once.
 * index_base is added to local indices in summary.best_index (shard offset). */
int rqcd_generate_check(rqcd_ctx* ctx, const rqcd_bc_soa* in, int64_t n, double min_time_section,
                        uint32_t flags, int64_t index_base, rqcd_out* out);

/* CollisionChecker(Trajectory).CollisionCheck(obstacle, minTimeSection) for pre-built quintics against
 * the whole set (CollisionChecker.hpp:40-61, CollisionChecker.cpp:70-193). `cost` (n doubles, may be NULL)
 * only feeds summary.best_*. */
int rqcd_check(rqcd_ctx* ctx, const rqcd_coeff_soa* in, const double* cost, int64_t n,
               double min_time_section, uint32_t flags, int64_t index_base, rqcd_out* out);

/* PerformanceEval shape (PerformanceEval.cpp:140-182): trajectory i against its own sphere i.
 * spheres: 4 rows (cx, cy, cz, radius) of n doubles at spheres + r*sphere_stride. Only out->verdict,
 * cost, coeffs, summary are used. */
int rqcd_generate_check_pairwise(rqcd_ctx* ctx, const rqcd_bc_soa* in, const double* spheres,
                                 int64_t sphere_stride, int64_t n, double min_time_section,
                                 uint32_t flags, int64_t index_base, rqcd_out* out);

/* ---- plane boundaries --------------------------------------------------------
 * CollisionChecker::CollisionCheck(Boundary) (src/CollisionChecker.cpp:25-68; the same test is
 * RapidTrajectoryGenerator::CheckPositionFeasibility, src/RapidTrajectoryGenerator.cpp:162-212) for every
 * (trajectory, plane) pair, with the DOCUMENTED intent: Collision (1) when (X(t) - point).normal <= 0 at t_start,
 * at t_end, or at a time inside the window where the velocity along the normal vanishes; else NoCollision (0).
 * The reference's own code does not implement that (its solver call overwrites the prepended end times and two
 * uninitialised slots are read, CollisionChecker.cpp:43-55), so parity for this entry is defined against the
 * oracle's restatement of the intent, not against the reference binary.
 * planes: HOST pointer, p records of 6 doubles {point xyz, normal xyz} (Boundary, ConvexObj.hpp:29-53); the normal is
 * normalised like the reference does (Vec3::GetUnitVector). verdict[n]: 1 if any plane is hit; verdict_matrix[n*p]
 * row-major per trajectory. Either may be NULL. Host or device pointers per flags. */
int rqcd_check_planes(rqcd_ctx* ctx, const rqcd_coeff_soa* in, int64_t n, const double* planes, int32_t p,
                      uint32_t flags, uint8_t* verdict, uint8_t* verdict_matrix);
int rqcd_generate_check_planes(rqcd_ctx* ctx, const rqcd_bc_soa* in, int64_t n, const double* planes, int32_t p,
                               uint32_t flags, uint8_t* verdict, uint8_t* verdict_matrix);

/* ---- input feasibility ---------------------------------------------------------
 * RapidTrajectoryGenerator::CheckInputFeasibility(fminAllowed, fmaxAllowed, wmaxAllowed, minTimeSection)
 * (src/RapidTrajectoryGenerator.cpp:74-160) after Generate, per candidate; result[i] is an
 * rqcd_input_feasibility. gravity: HOST pointer to 3 doubles (the constructor's 4th argument).
 * group (n int64, may be NULL; host or device per flags): consecutive candidates with equal group[i] are treated as
 * Generate calls on ONE generator object, which is what ForestPerformanceEval.cpp:177-209 does per batch: the
 * reference then keeps each axis' acceleration peak times from the first candidate of the group that needed them
 * (SingleAxisTrajectory.cpp:133-158 caches them and only Reset() clears the cache). NULL = a fresh object per
 * candidate (PerformanceEval.cpp:140-164). */
int rqcd_check_input_feasibility(rqcd_ctx* ctx, const rqcd_bc_soa* in, int64_t n, const double* gravity, double fmin,
                           double fmax, double wmax, double min_time_section, const int64_t* group, uint32_t flags,
                           uint8_t* result);

/* The same test as a FILTER inside the generate + check calls -- what both drivers do between Generate and CollisionCheck
 * (PerformanceEval.cpp:163-173 keeps only InputFeasible trials; a fresh generator object per candidate). While a filter is
 * set, rqcd_generate_check and rqcd_generate_check_pairwise first run CheckInputFeasibility(fmin, fmax, wmax,
 * min_time_section) on every candidate; a candidate that is not InputFeasible is NOT checked and NOT counted in the
 * summary, its verdict is RQCD_NOT_CHECKED, its first_obstacle -2, its row of the verdict matrix is not written.
 * rqcd_out::input_feasibility (optional) receives the per-candidate result. limits == NULL switches the filter off. */
typedef struct rqcd_input_limits {
  double gravity[3];
  double fmin, fmax, wmax;   /* CheckInputFeasibility(fminAllowed, fmaxAllowed, wmaxAllowed, .) */
  double min_time_section;
} rqcd_input_limits;
#define RQCD_NOT_CHECKED 255
int rqcd_set_input_filter(rqcd_ctx* ctx, const rqcd_input_limits* limits);

/* ---- the planner's candidate loop ----------------------------------------------
 * The reference ships the stages but not the loop that chains them (its docs describe one, StoppingTrajFinder::
 * GetBestTraj, whose sources are not in the tree): generate every candidate, drop those that cost too much, that are
 * not input feasible, that cross a plane boundary or that hit an obstacle, and return the cheapest survivor. This
 * entry point is that chain on the device; each stage is the corresponding call above, so per stage the results are
 * the reference's. stage[i] (may be NULL) says where candidate i stopped. */
typedef enum rqcd_plan_stage {
  RQCD_PLAN_FEASIBLE = 0,        /* survived every stage: candidates for summary.best_* */
  RQCD_PLAN_COST = 1,            /* GetCost() > max_cost (or NaN) */
  RQCD_PLAN_INPUT = 2,           /* CheckInputFeasibility != InputFeasible */
  RQCD_PLAN_BOUNDARY = 3,        /* CollisionCheck(Boundary) == Collision for some plane */
  RQCD_PLAN_COLLISION = 4,       /* CollisionCheck(obstacle) == Collision for some obstacle */
  RQCD_PLAN_INDETERMINABLE = 5   /* ... == CollisionIndeterminable */
} rqcd_plan_stage;

typedef struct rqcd_plan_params {
  double max_cost;          /* +inf: keep all */
  int32_t check_input;      /* 0: skip the input-feasibility stage */
  double gravity[3];
  double fmin, fmax, wmax;  /* CheckInputFeasibility(fminAllowed, fmaxAllowed, wmaxAllowed, .) */
  const double* planes;     /* HOST pointer, n_planes x 6 {point, normal}; may be NULL */
  int32_t n_planes;
  double min_time_section;  /* for both recursive tests */
} rqcd_plan_params;

/* in: boundary conditions (host or device per flags). stage: [n] bytes (host or device per flags) or NULL.
 * summary: HOST pointer, required; traj_counts / pair_counts cover the candidates that reached the obstacle stage,
 * best_cost / best_index the cheapest RQCD_PLAN_FEASIBLE candidate (index_base + i, lowest index on ties). */
int rqcd_plan_best(rqcd_ctx* ctx, const rqcd_bc_soa* in, int64_t n, const rqcd_plan_params* params, uint32_t flags,
                   int64_t index_base, uint8_t* stage, rqcd_summary* summary);

/* ---- solver unit entry points (test surface of Quartic::solveP3 / solve_quartic,
 * src/quartic.cpp:39-156). coef: 3 (cubic a,b,c) or 4 (quartic a,b,c,d) rows of n doubles, row stride n.
 * roots: 3 or 4 rows of n doubles (row stride n); count[n]. Host or device pointers per flags. */
int rqcd_solve_cubic(rqcd_ctx* ctx, const double* coef, int64_t n, uint32_t flags, double* roots, int32_t* count);
int rqcd_solve_quartic(rqcd_ctx* ctx, const double* coef, int64_t n, uint32_t flags, double* roots, int32_t* count);
/* How much each problem's answer depends on the last bits of the solver's acos / cos / pow -- the only operations of the
 * path that are not the reference's own IEEE sequence (see RQCD_F_FRAGILE). degree 3 or 4, coef as above.
 * cond[n]: the largest movement of any reported root when those outputs move by eta = 2^-50, divided by eta -- two
 * conforming math libraries (<= 2 ulp apart) give roots within about cond * 2^-51 of each other, whatever that is
 * relative to the root (the closed forms lose up to half of the digits on nearly repeated roots).
 * branch_fragile[n]: 1 when a branch of the solver (three or one real root of the resolvent reported, which resolvent
 * root, |D| < eps, the sign of a discriminant) is within that movement of its threshold: the COUNT is then libm's choice. */
int rqcd_solve_conditioning(rqcd_ctx* ctx, int32_t degree, const double* coef, int64_t n, uint32_t flags, double* cond,
                            uint8_t* branch_fragile);

/* ---- multi-GPU ---------------------------------------------------------------
 * Candidates are sharded by index into contiguous ranges, one per GPU; each GPU holds the full obstacle set. The only
 * exchange is the merge of the per-shard summaries (counts add; best = min cost, lowest index on ties), 88 bytes per
 * shard -- the reduction the drivers' loops do on the host (ForestPerformanceEval.cpp:216-230 counts verdicts; a planner
 * keeps the cheapest collision-free candidate). Two ways to run it, both inside this library, neither with a host round
 * trip between the check kernel and the merged record:
 *
 * (1) one process, several GPUs (a C++ host): rqcd_group_*. Each device's summary is stored by a one-thread kernel
 *     into its slot of a table in the FIRST device's memory through the NVLink peer mapping (mapped host memory when
 *     the devices cannot reach each other), and a merge kernel there applies the rule.
 * (2) one process per GPU (torchrun, MPI): rqcd_comm_* + rqcd_exchange_summaries. The context owns an NCCL
 *     communicator; the exchange is an ncclAllGather of the records on the context's stream followed by the same
 *     merge kernel on every rank. libnccl.so.2 is loaded at rqcd_comm_init time (dlopen; RQCD_NCCL_LIB overrides the
 *     name), so hosts that never call it do not need NCCL.
 *
 * rqcd_merge_summaries is the same rule as pure host arithmetic, for hosts that carry the records themselves. */
int rqcd_merge_summaries(const rqcd_summary* shards, int32_t n_shards, rqcd_summary* merged);

/* [lo, hi) of shard k of n candidates over n_shards: [k*n/n_shards, (k+1)*n/n_shards). */
void rqcd_shard_range(int64_t n, int32_t k, int32_t n_shards, int64_t* lo, int64_t* hi);

/* (1) device group. devices may name a device more than once (two shards on one GPU; used by the tests). */
typedef struct rqcd_group rqcd_group;
int rqcd_group_create(rqcd_group** group, const int32_t* devices, int32_t n_devices);
void rqcd_group_destroy(rqcd_group* group);
int32_t rqcd_group_size(const rqcd_group* group);
rqcd_ctx* rqcd_group_ctx(rqcd_group* group, int32_t k); /* shard k's context (streams, tuning, errors) */
const char* rqcd_group_last_error(const rqcd_group* group);
int rqcd_group_peer_slots(const rqcd_group* group);     /* 1: summary slots are peer-mapped device memory, 0: mapped host memory */
/* rqcd_set_obstacles on every device of the group. */
int rqcd_group_set_obstacles(rqcd_group* group, const rqcd_obstacle* obstacles, int32_t m);
/* rqcd_generate_check over the whole group with HOST buffers: shard k takes candidates rqcd_shard_range(n, k, size)
 * of `in` and writes the same range of every array in `out`; each shard runs rqcd_generate_check's streaming host
 * pipeline on its own device from its own host thread. out->summary (HOST, may be NULL) receives the merged record.
 * RQCD_F_DEVICE_PTRS is not allowed here. */
int rqcd_group_generate_check(rqcd_group* group, const rqcd_bc_soa* in, int64_t n, double min_time_section,
                              uint32_t flags, rqcd_out* out);
/* The same with shards already RESIDENT on their devices: in[k], n[k], out[k] (device pointers on device k, out may be
 * NULL, out[k].summary ignored) describe shard k; index_base of shard k is index_base + n[0] + .. + n[k-1]. The call
 * only enqueues (check, publish, merge); merged (HOST, may be NULL) makes it wait for and return the merged record,
 * otherwise rqcd_group_read_summary does. */
int rqcd_group_generate_check_resident(rqcd_group* group, const rqcd_bc_soa* in, const int64_t* n,
                                       double min_time_section, uint32_t flags, int64_t index_base,
                                       const rqcd_out* out, rqcd_summary* merged);
int rqcd_group_read_summary(rqcd_group* group, rqcd_summary* merged);

/* (2) one rank per process. The 128-byte id comes from rqcd_comm_unique_id on one rank and reaches the others over
 * the host's own transport (torch.distributed broadcast, MPI_Bcast, a file). */
#define RQCD_COMM_ID_BYTES 128
int rqcd_comm_unique_id(void* id);
int rqcd_comm_init(rqcd_ctx* ctx, const void* id, int32_t rank, int32_t world);
int rqcd_comm_finalize(rqcd_ctx* ctx);
int32_t rqcd_comm_world(const rqcd_ctx* ctx);  /* 1 without a communicator */
/* Enqueues, on the context's stream, the all-gather of the most recent check launch's summary and its merge; every
 * rank ends up with the same merged record in device memory. d_record (DEVICE pointer, may be NULL) receives a copy,
 * e.g. a slot of a per-step ring that is read once after many steps. Without a communicator (world 1) the merge is of
 * the one local record. No host synchronisation. */
int rqcd_exchange_summaries(rqcd_ctx* ctx, rqcd_summary* d_record);
/* Waits for the stream and returns the merged record of the last rqcd_exchange_summaries. */
int rqcd_read_merged_summary(rqcd_ctx* ctx, rqcd_summary* merged);

/* ---- synthetic inputs -------------------------------------------------------
 * Counter-based generator shared by host oracle and device (splitmix64 of (seed, global index, row)),
 * PerformanceEval's sampling ranges (PerformanceEval.cpp:59-76): fills a RQCD_BC_ROWS x n matrix at
 * d_bc (DEVICE pointer) for global indices [index_base, index_base+n). */
int rqcd_synth_bc(rqcd_ctx* ctx, uint64_t seed, int64_t index_base, int64_t n, double* d_bc, int64_t stride);

/* Measured FP64 issue ceilings of this GPU (for the roofline denominator): runs dependent-free
 * DFMA and DMUL+DADD loops for about `millis` ms each and returns flop/s, FMA counted as 2. */
int rqcd_measure_fp64_peak(rqcd_ctx* ctx, int millis, double* dfma_flops, double* dmul_dadd_flops);

#ifdef __cplusplus
}
#endif
#endif /* RQCD_H_ */
