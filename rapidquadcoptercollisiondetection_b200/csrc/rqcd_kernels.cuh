// rqcd_kernels.cuh -- launch-side declarations shared by rqcd_kernels.cu (device code, compiled with
// -fmad=false) and rqcd_api.cu (C ABI, host logic).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rqcd {

#ifndef RQCD_THREADS
#define RQCD_THREADS 256
#endif
#ifndef RQCD_MIN_BLOCKS
#define RQCD_MIN_BLOCKS 2
#endif
constexpr int kThreads = RQCD_THREADS;  // threads per CTA of the check kernel
constexpr int kMaxDepth = 40;     // RQCD_MAX_DEPTH: pending intervals per check

// One CTA's contribution to rqcd_summary; reduced by k_finalize in block order (deterministic).
struct BlockPartial {
  unsigned long long traj[4];
  unsigned long long pair[4];
  double best_cost;
  long long best_index;
};

struct DevSummary {  // mirrors rqcd_summary
  unsigned long long traj[4];
  unsigned long long pair[4];
  unsigned long long pairs_checked;
  double best_cost;
  long long best_index;
};

enum CheckMode { MODE_BC = 0, MODE_COEFF = 1 };

struct CheckParams {
  // trajectories
  const double* in;        // MODE_BC: 19 boundary-condition rows; MODE_COEFF: 18 coefficient rows
  long long in_stride;
  // MODE_BC, rqcd_bc_soa::shared_rows: bit r set = row r holds one value per GROUP of group_size consecutive candidates.
  // Candidate i of the launch reads element i + in_first of a per-candidate row and element (i + grp_first) / group_size
  // of a shared row (see bc_index in rqcd_kernel_util.cuh).
  unsigned shared_rows;
  long long in_first, grp_first, group_size;
  const double* t_start;   // MODE_COEFF, may be null (0)
  const double* t_end;     // MODE_COEFF
  const double* cost_in;   // MODE_COEFF, may be null (NaN)
  const double* spheres;   // pairwise: 4 rows cx, cy, cz, radius
  long long sphere_stride;
  unsigned goal_mask;
  long long n;
  long long index_base;
  double min_t;
  // obstacle table blob in global memory: mp records of `stride` doubles (stride odd), then mp int flags
  const void* obs_blob;
  int m, mp, stride;
  // outputs, each may be null
  uint8_t* verdict;
  int* first_obstacle;
  uint8_t* verdict_matrix;
  double* cost;
  double* coeffs;
  long long coeff_stride;
  // control
  int early_exit;
  int refill_min;          // idle lanes per warp that trigger a work-queue fetch
  int lockstep;            // 1: the warps of a CTA meet at a barrier every trip (see k_check)
  int th_solve, th_plane;  // k_pool: lanes waiting for the solver / plane stage that trigger that kind of trip
  unsigned long long* work_counter;
  BlockPartial* partials;  // [gridDim.x]
  // streaming input (host-pointer calls): candidates [0, *progress) have arrived in `in`; null = all present.
  // The copy engine advances *progress after each chunk; *stream_error is set if a wait times out.
  const unsigned long long* progress;
  unsigned long long* stream_error;
  long long stream_timeout;  // SM cycles a warp waits for a chunk before it fails the call
  // planner: candidates with skip[i] != 0 were rejected by an earlier stage and are not checked (nor counted)
  const uint8_t* skip;
  // input filter: the candidates to check are list[0 .. *n_list) (indices into the launch's rows, any order; written by
  // k_feas just before); null = all of [0, n)
  const long long* list;
  const unsigned long long* n_list;
};

// once per device, before any kernel that solves a cubic: the refined reciprocals of solveP3's constant divisors
// (one copy per kernel translation unit)
cudaError_t init_device_constants(cudaStream_t s);
cudaError_t init_pool_constants(cudaStream_t s);

// ---- k_check (rqcd_kernels.cu): lane owns trajectory; every shape
size_t check_smem_bytes(int mp, int stride, bool moving);
cudaError_t launch_check(const CheckParams& p, int mode, bool moving, bool pairwise, int grid, cudaStream_t s);
int check_max_blocks_per_sm(int mode, bool moving, bool pairwise, int m, size_t smem);
cudaError_t check_prepare(int mode, bool moving, bool pairwise, int m, size_t smem);  // opt in to large dynamic smem
const char* check_kernel_name(bool moving, bool pairwise, int m);
constexpr int kEndBitsMinObstacles = 16;  // static tables at least this long: end-point tests as a pre-pass
cudaError_t launch_finalize(const BlockPartial* partials, int grid, DevSummary* d_summary, cudaStream_t s);
// RQCD_F_FRAGILE (k_fragile): every pair's recursion once more, in lockstep with eight copies whose cubic-solver
// transcendentals are moved by +-eta; a pair with a sign decision that is not `margin` times farther from zero than the
// copies scatter is marked: bit 7 of matrix[i*m + k], and fragile[i] when the pair is at or before the trajectory's
// first hit first[i] (null: every pair counts). Null arrays are skipped.
cudaError_t launch_fragile(const CheckParams& p, int mode, bool moving, bool pairwise, double eta, double margin, double floor_ulps,
                           const int* first, uint8_t* fragile, uint8_t* matrix, unsigned long long* why /* [32] or null */,
                           cudaStream_t s);
// multi-GPU exchange (rqcd_multi.cu): a shard's summary to its slot of a table (peer-mapped memory of the group's first
// device, or the all-gather's receive buffer); the merge of all slots into *out (and *copy when given)
cudaError_t launch_publish_summary(const DevSummary* mine, DevSummary* slot, cudaStream_t s);
cudaError_t launch_merge_summaries(const DevSummary* slots, int n, DevSummary* out, DevSummary* copy, cudaStream_t s);

// ---- k_pool (rqcd_pool.cu): staged sections with the certified solver skip; obstacle tables, static or moving
size_t pool_smem_bytes(int mp, int stride, int threads);
int pool_threads(int mp, int stride, size_t smem_per_sm, size_t smem_per_block_max);  // 256 / 384 / 512, 0 = does not fit
cudaError_t pool_prepare(int mode, bool moving, int threads, size_t smem);
int pool_max_blocks_per_sm(int mode, bool moving, int threads, size_t smem);
cudaError_t launch_pool(const CheckParams& p, int mode, bool moving, int threads, int grid, size_t smem, cudaStream_t s);
const char* pool_kernel_name(bool moving);

cudaError_t launch_planes(const CheckParams& p, int mode, const double* planes, int np, uint8_t* verdict, uint8_t* matrix,
                          cudaStream_t s);
// group may be null; peaks: scratch of 6*n doubles + n bytes, needed when group is given
cudaError_t launch_feasibility(const CheckParams& p, const double grav[3], double fmin, double fmax, double wmax,
                               double min_t, const long long* group, double* peaks, uint8_t* result, uint8_t* verdict,
                               int* first, long long* list /* accepted candidates, ascending; may be null */,
                               unsigned long long* n_list /* count, followed by feasibility_list_words(n) - 1 words of scratch */,
                               cudaStream_t s);
size_t feasibility_list_words(long long n);
// planner glue (rqcd_plan_best): stage codes, see rqcd_plan_stage in rqcd.h
cudaError_t launch_plan_mask(long long n, const double* cost, double max_cost, const uint8_t* feas, const uint8_t* plane_hit,
                             uint8_t* stage, cudaStream_t s);
cudaError_t launch_plan_stage(long long n, const uint8_t* verdict, uint8_t* stage, cudaStream_t s);
cudaError_t launch_generate(const CheckParams& p, double* cost, double* coeffs, long long coeff_stride, double* params,
                            long long param_stride, cudaStream_t s);
cudaError_t launch_solve(bool quartic, const double* coef, long long n, double* roots, int* count, cudaStream_t s);
cudaError_t launch_solve_cond(bool quartic, const double* coef, long long n, double eta, double margin, double floor_ulps,
                              double* cond, uint8_t* flag, cudaStream_t s);
cudaError_t launch_synth_bc(unsigned long long seed, long long index_base, long long n, double* bc, long long stride,
                            cudaStream_t s);
cudaError_t launch_fp64_peak(bool fma, int blocks, int iters, double* sink, cudaStream_t s);
constexpr int kPeakThreads = 256;
constexpr int kPeakChains = 8;      // independent accumulators per thread
constexpr int kPeakInnerOps = 64;   // operations per chain per outer iteration

}  // namespace rqcd
