/*
 * oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C99) of the reference algorithm on the hot path of
 * nlbucki/RapidQuadcopterCollisionDetection.  It exists to check the CUDA path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it; the product library (librqcd.so) never links or calls it.
 *
 * Parity status: PINNED (one exception: orc_check_planes, see below).  oracle.c is checked (tests/test_oracle_*.py) against
 *  - the reference's checked-in known answers (data/PerformanceEval.txt:23-25,
 *    data/ForestPerfEval.txt:28-29, data/ForestPerfEval.csv, data/ForestPerfEvaltreeParamsCSV.csv), and
 *  - oracle/_ref/libref.so, the UNMODIFIED reference sources compiled where they lie
 *    (oracle/Makefile), bit-for-bit on verdicts, coefficients, costs and roots.
 */
#ifndef RQCD_ORACLE_H_
#define RQCD_ORACLE_H_

#include "../include/rqcd.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Work counters for the flop model of SURVEY.md section 8(d). */
typedef struct orc_stats {
  uint64_t checks;           /* CollisionCheck(obstacle) calls */
  uint64_t sections;         /* S  : CollisionCheckSection entries */
  uint64_t tangent_sections; /* S_t: sections that reach GetTangentPlane */
  uint64_t test_points;      /* P  : plane-side test points evaluated */
  uint64_t trig;             /* solveP3 three-real-root branch */
  uint64_t cardano;          /* solveP3 one-real-root branch */
  uint64_t root_pairs;       /* real root pairs appended by solve_quartic */
  uint64_t cubic_fallback;   /* |c0| <= 1e-6 branch (CollisionChecker.cpp:119-120) */
  uint64_t max_depth;        /* deepest recursion seen */
  uint64_t sphere_checks, prism_checks;
  /* statistics for the certified solver skip of the CUDA path: tangent sections whose scalar quintic
   * D(t) = (X(t) - point).normal has all six Bernstein coefficients on [ts, tf] positive -- there every plane-side test
   * of the reference is negative whatever the roots are, so the section is NoCollision without the root solver. Counted
   * apart for a pair's first section and for deeper ones. Statistics only: no verdict depends on them. */
  uint64_t skip_first, skip_child, child_tangent_sections;
} orc_stats;

/* Quartic::solveP3 / solve_quartic (src/quartic.cpp:39-72, 77-156). */
int orc_solve_p3(double a, double b, double c, double* x);
int orc_solve_quartic(double a, double b, double c, double d, double* root);
void orc_solve_cubic_batch(const double* coef, int64_t n, double* roots, int32_t* count);
void orc_solve_quartic_batch(const double* coef, int64_t n, double* roots, int32_t* count);

/* One CollisionChecker(traj).CollisionCheck(obstacle, minT) (src/CollisionChecker.cpp:70-193). */
int orc_check_one(const double coef[18], double t_start, double t_end, const rqcd_obstacle* obs,
                  double min_time_section, orc_stats* stats);

/* Batched mirrors of the C ABI in include/rqcd.h (same structs, HOST pointers only). */
void orc_generate(const rqcd_bc_soa* in, int64_t n, rqcd_out* out, double* abg /* 9 rows x n or NULL */);
void orc_generate_check(const rqcd_bc_soa* in, int64_t n, const rqcd_obstacle* obs, int32_t m,
                        double min_time_section, uint32_t flags, int64_t index_base, rqcd_out* out,
                        int nthreads, orc_stats* stats);
void orc_check(const rqcd_coeff_soa* in, const double* cost, int64_t n, const rqcd_obstacle* obs, int32_t m,
               double min_time_section, uint32_t flags, int64_t index_base, rqcd_out* out, int nthreads,
               orc_stats* stats);
void orc_generate_check_pairwise(const rqcd_bc_soa* in, const double* spheres, int64_t sphere_stride,
                                 int64_t n, double min_time_section, int64_t index_base, rqcd_out* out,
                                 int nthreads, orc_stats* stats);

/* CollisionChecker::CollisionCheck(Boundary) (src/CollisionChecker.cpp:25-68), the documented intent; parity
 * UNPINNED against the reference binary (its code reads uninitialised memory there, see oracle.c). */
int orc_check_plane_one(const double coef[18], double t0, double t1, const double plane[6]);
void orc_check_planes(const rqcd_coeff_soa* in, int64_t n, const double* planes, int32_t p, uint8_t* verdict,
                      uint8_t* matrix);

/* RapidTrajectoryGenerator::CheckInputFeasibility (src/RapidTrajectoryGenerator.cpp:74-160).
 * group (or NULL): consecutive trajectories with equal group[i] are checked with ONE generator object,
 * as the Forest driver does per batch (ForestPerformanceEval.cpp:177-190), so the per-axis cache
 * _accPeakTimes (SingleAxisTrajectory.cpp:133-158, cleared only by Reset) keeps the peak times of the
 * first candidate of the group that reached GetMinMaxAcc on that axis.  NULL = fresh object each. */
void orc_input_feasibility(const rqcd_bc_soa* in, int64_t n, const double gravity[3], double fmin, double fmax,
                           double wmax, double min_time_section, const int64_t* group, uint8_t* result);

/* splitmix64 counter-based synthetic boundary conditions (same function as rqcd_synth_bc). */
/* sections entered, per-axis bound evaluations, sections that reached the rate bound -- since the last reset */
void orc_feas_counters(uint64_t out[3], int reset);
void orc_synth_bc(uint64_t seed, int64_t index_base, int64_t n, double* bc, int64_t stride);

/* The reference drivers' input streams: mt19937(0) + libstdc++ uniform_real_distribution<double>.
 * orc_c1_inputs keeps drawing until n_accept trials pass CheckInputFeasibility(5,30,20,0.002)
 * (PerformanceEval.cpp:117-173); returns total trials. orc_c2_inputs fills n_batches*per_batch candidates
 * (ForestPerformanceEval.cpp:167-193). */
int64_t orc_c1_inputs(int64_t n_accept, int filter, double* bc, int64_t stride, double* spheres,
                      int64_t sphere_stride);
void orc_c2_inputs(int64_t n_batches, int64_t per_batch, double* bc, int64_t stride);
/* The five Forest prisms (ForestPerformanceEval.cpp:130-136). */
void orc_forest_obstacles(rqcd_obstacle out[5]);

/* Rotation::GetRotationMatrix (include/CommonMath/Rotation.hpp:203-220). */
void orc_rotation_matrix(const double quat[4], double R[9]);
/* Trajectory::GetValue (include/CommonMath/Trajectory.hpp:76-82). */
void orc_traj_value(const double coef[18], double t, double out[3]);

#ifdef __cplusplus
}
#endif
#endif
