// rqcd_api.cu -- the C ABI of include/rqcd.h: context, obstacle upload, staging for host buffers and the
// launches. Host logic only; all arithmetic of the path lives in rqcd_kernels.cu / rqcd_device.cuh, except
// the per-obstacle constants hoisted in prepare_obstacle() below.
//
// There is no CPU implementation of the path in this library: without a CUDA device rqcd_create fails.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <vector>

#include "../../include/rqcd.h"
#include "rqcd_device.cuh"
#include "rqcd_host.cuh"
#include "rqcd_kernels.cuh"

using namespace rqcd;

static_assert(RQCD_MAX_DEPTH == kMaxDepth, "RQCD_MAX_DEPTH and the device stack size must agree");
static_assert(sizeof(DevSummary) == sizeof(rqcd_summary), "DevSummary mirrors rqcd_summary");

namespace {

cudaError_t ensure(DevBuf& b, size_t bytes) {
  if (bytes <= b.cap) return cudaSuccess;
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
  size_t want = bytes + bytes / 8;
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) return e;
  b.cap = want;
  return cudaSuccess;
}

void release(DevBuf& b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
}

bool finite3(const double* v) { return std::isfinite(v[0]) && std::isfinite(v[1]) && std::isfinite(v[2]); }

// Rotation::GetRotationMatrix (include/CommonMath/Rotation.hpp:203-220), same expression order, so that
// hoisting it out of Rotation::Rotate (Rotation.hpp:224-233) is bit-identical to rebuilding it per call.
void rotation_matrix(const double v[4], double R[9]) {
  const double r0 = v[0] * v[0];
  const double r1 = v[1] * v[1];
  const double r2 = v[2] * v[2];
  const double r3 = v[3] * v[3];
  R[0] = r0 + r1 - r2 - r3;
  R[1] = 2 * v[1] * v[2] - 2 * v[0] * v[3];
  R[2] = 2 * v[1] * v[3] + 2 * v[0] * v[2];
  R[3] = 2 * v[1] * v[2] + 2 * v[0] * v[3];
  R[4] = r0 - r1 + r2 - r3;
  R[5] = 2 * v[2] * v[3] - 2 * v[0] * v[1];
  R[6] = 2 * v[1] * v[3] - 2 * v[0] * v[2];
  R[7] = 2 * v[2] * v[3] + 2 * v[0] * v[1];
  R[8] = r0 - r1 - r2 + r3;
}

int validate_obstacle(const rqcd_obstacle& o) {
  if (o.type != RQCD_SPHERE && o.type != RQCD_RECT_PRISM) return RQCD_ERR_UNSUPPORTED;
  if (o.moving != 0 && o.moving != 1) return RQCD_ERR_INVALID_ARG;
  if (!finite3(o.center)) return RQCD_ERR_INVALID_ARG;
  if (o.type == RQCD_SPHERE) {
    if (!(o.radius > 0) || !std::isfinite(o.radius)) return RQCD_ERR_INVALID_ARG;  // Sphere.hpp:44
  } else {
    if (!(o.side[0] > 0 && o.side[1] > 0 && o.side[2] > 0) || !finite3(o.side)) return RQCD_ERR_INVALID_ARG;  // RectPrism.hpp:52
    for (int i = 0; i < 4; i++)
      if (!std::isfinite(o.quat[i])) return RQCD_ERR_INVALID_ARG;
  }
  if (o.moving) {
    for (int i = 0; i < 18; i++)
      if (!std::isfinite(o.motion[i])) return RQCD_ERR_INVALID_ARG;
    if (!(o.motion_t_start <= o.motion_t_end)) return RQCD_ERR_INVALID_ARG;  // Trajectory.hpp:68
  }
  return RQCD_OK;
}

int ensure_partials(rqcd_ctx* ctx, int grid) {
  if (grid <= ctx->partial_cap) return RQCD_OK;
  if (ctx->d_partials) cudaFree(ctx->d_partials);
  ctx->d_partials = nullptr;
  ctx->partial_cap = 0;
  RQCD_CU(cudaMalloc(&ctx->d_partials, sizeof(BlockPartial) * (size_t)grid));
  ctx->partial_cap = grid;
  return RQCD_OK;
}

// Everything below `what the caller passed` is device memory here.
struct RowBlock {  // host rows still to be copied: row r goes to dst + r*n (device), from src + r*src_stride (host)
  double* dst;
  const double* src;
  long long src_stride;
  int rows;
  long long count;  // < 0: one element per candidate (copied chunk by chunk); else elements per row, copied before chunk 0
};

constexpr int kMaxBlocks = RQCD_BC_ROWS + 4;

struct BlockList {
  RowBlock blocks[kMaxBlocks];
  int nblocks = 0;
  void add(void* dst, const double* src, long long src_stride, int rows, long long count = -1) {
    blocks[nblocks++] = RowBlock{(double*)dst, src, src_stride, rows, count};
  }
};

}  // namespace

// RQCD_F_FAST_FMA: the kernel files' second build (csrc/Makefile: -fmad=true, namespace rqcd_fma). Its parameter structs
// are the same definitions under another namespace name, hence the same layout.
namespace rqcd_fma {
struct CheckParams;
cudaError_t init_device_constants(cudaStream_t s);
cudaError_t init_pool_constants(cudaStream_t s);
cudaError_t launch_check(const CheckParams& p, int mode, bool moving, bool pairwise, int grid, cudaStream_t s);
int check_max_blocks_per_sm(int mode, bool moving, bool pairwise, int m, size_t smem);
cudaError_t check_prepare(int mode, bool moving, bool pairwise, int m, size_t smem);
cudaError_t pool_prepare(int mode, bool moving, int threads, size_t smem);
int pool_max_blocks_per_sm(int mode, bool moving, int threads, size_t smem);
cudaError_t launch_pool(const CheckParams& p, int mode, bool moving, int threads, int grid, size_t smem, cudaStream_t s);
}  // namespace rqcd_fma

namespace {

struct CheckCall : BlockList {
  int mode = MODE_BC;
  bool pairwise = false;
  CheckParams p = {};
  bool fast_fma = false;  // RQCD_F_FAST_FMA
  const uint8_t* filter_mask = nullptr;  // input filter: != 0 where the candidate was rejected (k_fragile skips those)
};

// ---- rqcd_bc_soa ------------------------------------------------------------------------------------------------------
bool bad_bc(const rqcd_bc_soa* in, long long n) {
  if (!in || n < 0 || in->first < 0) return true;
  if (n > 0 && !in->data) return true;
  if ((in->shared_rows >> RQCD_BC_ROWS) != 0u) return true;
  if (in->shared_rows != 0u && in->group_size < 1) return true;
  return in->stride < in->first + n;
}

// device pointers: the kernel reads the caller's matrix in place
void bind_bc(CheckParams& p, const rqcd_bc_soa* in) {
  p.in = in->data;
  p.in_stride = in->stride;
  p.goal_mask = in->goal_mask;
  p.shared_rows = in->shared_rows;
  p.in_first = in->first;
  p.grp_first = in->first;
  p.group_size = in->shared_rows ? in->group_size : 1;
}

// host pointers: the rows of candidates [first, first + n) go to a staging matrix of RQCD_BC_ROWS x n doubles; of a
// shared row only the groups those candidates belong to cross PCIe. Runs of rows of one kind are one pitched copy.
void stage_bc(CheckParams& p, BlockList& bl, const rqcd_bc_soa* in, long long n, void* staging) {
  double* dst = (double*)staging;
  const long long G = in->shared_rows ? in->group_size : 1;
  const long long g0 = in->first / G;
  const long long ng = n > 0 ? (in->first + n - 1) / G - g0 + 1 : 0;
  for (int r = 0; r < RQCD_BC_ROWS;) {
    const unsigned kind = (in->shared_rows >> r) & 1u;
    int e = r + 1;
    while (e < RQCD_BC_ROWS && ((in->shared_rows >> e) & 1u) == kind) e++;
    if (kind) bl.add(dst + (long long)r * n, in->data + (long long)r * in->stride + g0, in->stride, e - r, ng);
    else bl.add(dst + (long long)r * n, in->data + (long long)r * in->stride + in->first, in->stride, e - r);
    r = e;
  }
  p.in = dst;
  p.in_stride = n;
  p.goal_mask = in->goal_mask;
  p.shared_rows = in->shared_rows;
  p.in_first = 0;
  p.grp_first = in->first - g0 * G;
  p.group_size = G;
}

// Which kernel a check call runs on: the CTA size of k_pool (staged sections) for obstacle tables that fit beside the
// trajectory rows, 0 = k_check (the pairwise shape, empty and oversized tables, short static tables).
int pool_threads_for(const rqcd_ctx* ctx, bool pairwise) {
  const bool moving = !pairwise && ctx->any_moving;
  // short static tables (the Forest's five trees: 4.4 executed pairs per candidate with early exit) spend their time
  // building candidates, not in sections: k_check's lane-owns-trajectory loop is 30 % faster there (5.2e9 vs 3.9e9 checks/s)
  const bool short_static = !moving && ctx->m < kEndBitsMinObstacles && ctx->pool < 2;
  if (ctx->pool && !short_static && !pairwise && ctx->m >= 1 && ctx->m < 65536) {
    const int t = pool_threads(ctx->mp, ctx->stride, ctx->prop.sharedMemPerMultiprocessor, kMaxSmem);
    // RQCD_POOL_THREADS (256 / 384 / 512): the CTA size by hand, when that size fits
    if (t > 0 && ctx->pool_threads_override > 0 && pool_smem_bytes(ctx->mp, ctx->stride, ctx->pool_threads_override) <= kMaxSmem)
      return ctx->pool_threads_override;
    return t;
  }
  return 0;
}

int run_check(rqcd_ctx* ctx, CheckCall& cc) {
  const bool moving = !cc.pairwise && ctx->any_moving;
  CheckParams& p = cc.p;
  p.obs_blob = ctx->obs_blob.p;
  p.m = cc.pairwise ? 1 : ctx->m;
  p.mp = ctx->mp;
  p.stride = ctx->stride;
  p.work_counter = ctx->d_counter;
  p.stream_timeout = ctx->stream_timeout;
  const int sms = ctx->prop.multiProcessorCount;
  // ---- which kernel: obstacle tables go to k_pool (staged sections with the certified solver skip) when the table
  // fits beside the trajectory rows; the pairwise shape, empty tables and oversized tables go to k_check
  const int pool_t = pool_threads_for(ctx, cc.pairwise);
  long long grid;
  if (pool_t > 0) {
    p.th_solve = ctx->th_solve;
    p.th_plane = ctx->th_plane;
    const size_t smem = pool_smem_bytes(ctx->mp, ctx->stride, pool_t);
    RQCD_CU(cc.fast_fma ? rqcd_fma::pool_prepare(cc.mode, moving, pool_t, smem) : pool_prepare(cc.mode, moving, pool_t, smem));
    int per_sm = cc.fast_fma ? rqcd_fma::pool_max_blocks_per_sm(cc.mode, moving, pool_t, smem)
                             : pool_max_blocks_per_sm(cc.mode, moving, pool_t, smem);
    if (per_sm < 1) per_sm = 1;
    grid = (long long)sms * per_sm;
    const long long need = (p.n + pool_t - 1) / pool_t;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    int st = ensure_partials(ctx, (int)grid);
    if (st != RQCD_OK) return st;
    p.partials = ctx->d_partials;
    RQCD_CU(cudaMemsetAsync(p.work_counter, 0, sizeof(unsigned long long), ctx->stream));
    if (cc.fast_fma)
      RQCD_CU(rqcd_fma::launch_pool(reinterpret_cast<const rqcd_fma::CheckParams&>(p), cc.mode, moving, pool_t, (int)grid, smem,
                                    ctx->stream));
    else
      RQCD_CU(launch_pool(p, cc.mode, moving, pool_t, (int)grid, smem, ctx->stream));
    ctx->last_kernel = pool_kernel_name(moving);
  } else {
    // CTA lockstep or free-running warps, and the idle lanes per warp that trigger a work-queue fetch (in lockstep:
    // the CTA's average): a fetch costs a pass in which only the fetching lanes (and, for a few of them, the helpers
    // that build their axes) work, an idle lane costs a slot in every frame until then. Measured optima on B200:
    // tables of 16+ obstacles run free and fetch at 6 idle lanes (4 from 128 obstacles on); short tables and the
    // pairwise shape, whose warps fetch every other trip, stay in lockstep (which also aligns their fetches) at 4 and 8.
    p.lockstep = ctx->lockstep >= 0 ? ctx->lockstep : ((!cc.pairwise && p.m >= 16) ? 0 : 1);
    p.refill_min = ctx->refill_min > 0 ? ctx->refill_min : (cc.pairwise ? 8 : (p.m >= 128 ? 4 : (p.m >= 16 ? 6 : 4)));
    const size_t smem = check_smem_bytes(cc.pairwise ? 0 : ctx->mp, cc.pairwise ? 0 : ctx->stride, moving);
    if (smem > kMaxSmem) return RQCD_ERR_UNSUPPORTED;
    RQCD_CU(cc.fast_fma ? rqcd_fma::check_prepare(cc.mode, moving, cc.pairwise, p.m, smem)
                        : check_prepare(cc.mode, moving, cc.pairwise, p.m, smem));
    int per_sm = cc.fast_fma ? rqcd_fma::check_max_blocks_per_sm(cc.mode, moving, cc.pairwise, p.m, smem)
                             : check_max_blocks_per_sm(cc.mode, moving, cc.pairwise, p.m, smem);
    if (per_sm < 1) per_sm = 1;
    grid = (long long)sms * per_sm;
    const long long need = (p.n + kThreads - 1) / kThreads;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    int st = ensure_partials(ctx, (int)grid);
    if (st != RQCD_OK) return st;
    p.partials = ctx->d_partials;
    RQCD_CU(cudaMemsetAsync(p.work_counter, 0, sizeof(unsigned long long), ctx->stream));
    if (cc.fast_fma)
      RQCD_CU(rqcd_fma::launch_check(reinterpret_cast<const rqcd_fma::CheckParams&>(p), cc.mode, moving, cc.pairwise, (int)grid,
                                     ctx->stream));
    else
      RQCD_CU(launch_check(p, cc.mode, moving, cc.pairwise, (int)grid, ctx->stream));
    ctx->last_kernel = check_kernel_name(moving, cc.pairwise, p.m);
  }
  RQCD_CU(launch_finalize(ctx->d_partials, (int)grid, ctx->d_summary, ctx->stream));
  ctx->launches += 2;  // the check kernel + k_finalize
  ctx->summary_valid = true;
  return RQCD_OK;
}

int fetch_summary(rqcd_ctx* ctx, rqcd_summary* out) {
  RQCD_CU(cudaMemcpyAsync(ctx->h_summary, ctx->d_summary, sizeof(DevSummary), cudaMemcpyDeviceToHost, ctx->stream));
  RQCD_CU(cudaStreamSynchronize(ctx->stream));
  memcpy(out, ctx->h_summary, sizeof(rqcd_summary));
  if (out->traj_counts[3] != 0 || out->pair_counts[3] != 0) return RQCD_ERR_DEPTH_CAP;
  return RQCD_OK;
}

bool bad_min_t(double t) { return !(t > 0) || !std::isfinite(t); }

// Host <-> device staging of row matrices (row r at base + r*stride, n doubles each).
cudaError_t rows_d2h(double* dst, long long stride, const void* src, long long n, int rows, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  return cudaMemcpy2DAsync(dst, (size_t)stride * 8, src, (size_t)n * 8, (size_t)n * 8, rows, cudaMemcpyDeviceToHost, s);
}

// every block in full (the calls that do not stream their input beside the kernel)
cudaError_t copy_blocks(const BlockList& bl, long long n, cudaStream_t s) {
  for (int b = 0; b < bl.nblocks; b++) {
    const RowBlock& rb = bl.blocks[b];
    const long long w = rb.count < 0 ? n : rb.count;
    if (w == 0) continue;
    cudaError_t e = cudaMemcpy2DAsync(rb.dst, (size_t)n * 8, rb.src, (size_t)rb.src_stride * 8, (size_t)w * 8, rb.rows,
                                      cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

struct OutStage {  // device-side output pointers + what must be copied back
  uint8_t* verdict = nullptr;
  int* first = nullptr;
  uint8_t* matrix = nullptr;
  double* cost = nullptr;
  double* coeffs = nullptr;
  long long coeff_stride = 0;
  // the per-candidate results go STRAIGHT to the caller's page-locked host buffer (the kernel stores through its device
  // alias: coalesced 32 / 128 / 256-byte posted writes over PCIe that overlap the rest of the kernel) instead of to a
  // staging buffer that is copied back after the kernel
  bool direct_verdict = false, direct_first = false, direct_cost = false;
};

// The device alias of a page-locked, mapped host buffer (cudaHostAlloc / cudaHostRegister under unified addressing), else null.
template <class T>
T* mapped_alias(T* host) {
  cudaPointerAttributes at;
  if (!host || cudaPointerGetAttributes(&at, host) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return (at.type == cudaMemoryTypeHost && at.devicePointer) ? (T*)at.devicePointer : nullptr;
}

int stage_outputs(rqcd_ctx* ctx, const rqcd_out* out, long long n, int m, bool want_cost_coeffs, OutStage& o, bool allow_direct) {
  const bool direct = allow_direct && ctx->direct_out;
  if (out->verdict) {
    uint8_t* alias = direct ? mapped_alias(out->verdict) : nullptr;
    if (alias) {
      o.verdict = alias, o.direct_verdict = true;
    } else {
      RQCD_CU(ensure(ctx->s_verdict, (size_t)n));
      o.verdict = (uint8_t*)ctx->s_verdict.p;
    }
  }
  if (out->first_obstacle) {
    int* alias = direct ? mapped_alias(out->first_obstacle) : nullptr;
    if (alias) {
      o.first = alias, o.direct_first = true;
    } else {
      RQCD_CU(ensure(ctx->s_first, (size_t)n * 4));
      o.first = (int*)ctx->s_first.p;
    }
  }
  if (out->verdict_matrix) {
    RQCD_CU(ensure(ctx->s_matrix, (size_t)n * (size_t)(m > 0 ? m : 1)));
    o.matrix = (uint8_t*)ctx->s_matrix.p;
  }
  if (want_cost_coeffs && out->cost) {
    double* alias = direct ? mapped_alias(out->cost) : nullptr;
    if (alias) {
      o.cost = alias, o.direct_cost = true;
    } else {
      RQCD_CU(ensure(ctx->s_cost, (size_t)n * 8));
      o.cost = (double*)ctx->s_cost.p;
    }
  }
  if (want_cost_coeffs && out->coeffs) {
    RQCD_CU(ensure(ctx->s_coeffs, (size_t)n * 8 * RQCD_COEFF_ROWS));
    o.coeffs = (double*)ctx->s_coeffs.p;
    o.coeff_stride = n;
  }
  return RQCD_OK;
}

void fill_params_out(CheckParams& p, const OutStage& o) {
  p.verdict = o.verdict;
  p.first_obstacle = o.first;
  p.verdict_matrix = o.matrix;
  p.cost = o.cost;
  p.coeffs = o.coeffs;
  p.coeff_stride = o.coeff_stride;
}

int validate_out(const rqcd_out* out, uint32_t flags, long long n) {
  if (!out) return RQCD_ERR_INVALID_ARG;
  if ((flags & RQCD_F_EARLY_EXIT) && out->verdict_matrix) return RQCD_ERR_INVALID_ARG;
  if (out->coeffs && out->coeff_stride < n) return RQCD_ERR_INVALID_ARG;
  if ((flags & RQCD_F_FRAGILE) && !out->fragile) return RQCD_ERR_INVALID_ARG;
  return RQCD_OK;
}

int outputs_d2h(rqcd_ctx* ctx, const rqcd_out* out, long long n, int m, const OutStage& o) {
  cudaStream_t cs = ctx->stream;
  if (n > 0) {
    if (o.verdict && out->verdict && !o.direct_verdict)
      RQCD_CU(cudaMemcpyAsync(out->verdict, o.verdict, (size_t)n, cudaMemcpyDeviceToHost, cs));
    if (o.first && out->first_obstacle && !o.direct_first)
      RQCD_CU(cudaMemcpyAsync(out->first_obstacle, o.first, (size_t)n * 4, cudaMemcpyDeviceToHost, cs));
    if (o.matrix && m > 0)
      RQCD_CU(cudaMemcpyAsync(out->verdict_matrix, o.matrix, (size_t)n * (size_t)m, cudaMemcpyDeviceToHost, cs));
    if (o.cost && !o.direct_cost) RQCD_CU(cudaMemcpyAsync(out->cost, o.cost, (size_t)n * 8, cudaMemcpyDeviceToHost, cs));
    if (o.coeffs) RQCD_CU(rows_d2h(out->coeffs, out->coeff_stride, o.coeffs, n, RQCD_COEFF_ROWS, cs));
  }
  return RQCD_OK;
}

// The input filter of the generate + check calls (rqcd_set_input_filter): CheckInputFeasibility for every candidate, the
// step both drivers take between Generate and CollisionCheck (PerformanceEval.cpp:163-173). The result doubles as the
// check kernels' skip mask: a candidate that is not InputFeasible is neither checked nor counted.
int run_input_filter(rqcd_ctx* ctx, CheckCall& cc, const OutStage& o, long long n, uint8_t* d_feas) {
  const rqcd_input_limits& L = ctx->filter;
  CheckParams q = cc.p;
  q.progress = nullptr;
  q.work_counter = ctx->d_counter;
  // the accepted candidates as a dense ascending list: the check kernel then has no empty lanes for rejected ones
  const size_t words = feasibility_list_words(n);
  RQCD_CU(ensure(ctx->s_list, ((size_t)n + words) * 8));
  unsigned long long* n_list = (unsigned long long*)ctx->s_list.p;
  long long* list = (long long*)ctx->s_list.p + words;
  RQCD_CU(launch_feasibility(q, L.gravity, L.fmin, L.fmax, L.wmax, L.min_time_section, nullptr, nullptr, d_feas, o.verdict,
                             cc.pairwise ? nullptr : o.first, list, n_list, ctx->stream));
  ctx->launches += 4;  // k_feas + the list's three
  cc.p.list = list;
  cc.p.n_list = n_list;
  cc.filter_mask = d_feas;
  return RQCD_OK;
}

// Host-pointer call: ONE persistent launch over the whole range that starts as soon as the first chunk of
// boundary conditions has crossed PCIe. The remaining chunks follow on the copy stream while the kernel runs; after
// each chunk the copy engine advances a progress word in device memory (an 8-byte copy from pinned memory -- no SM
// is involved, so a kernel that fills every SM cannot block it) and a warp that takes candidates beyond the
// progress mark waits for it (bounded: a wait that times out fails the call). Results return after the kernel.
int run_host_pipeline(rqcd_ctx* ctx, CheckCall& cc, const rqcd_out* out, long long n, int m, const OutStage& o) {
  // Streaming needs the copies to run on the copy engines while the kernel holds every SM. That is the case for
  // page-locked host memory; a pitched copy from PAGEABLE memory may be finished by a driver kernel, which could not
  // be scheduled next to the waiting persistent kernel -- so pageable input is copied completely before the launch.
  bool pinned = true;
  for (int b = 0; b < cc.nblocks; b++) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, cc.blocks[b].src) != cudaSuccess || at.type != cudaMemoryTypeHost) pinned = false;
  }
  cudaGetLastError();
  long long chunk = ctx->chunk_override > 0 ? ctx->chunk_override : (n + 31) / 32;
  if (chunk < 16384) chunk = 16384;  // the kernel can start once the first chunk is in: keep it small
  if (!pinned) chunk = n > 0 ? n : 1;
  if ((n + chunk - 1) / chunk > kMaxChunks) chunk = (n + kMaxChunks - 1) / kMaxChunks;
  chunk = (chunk + 255) / 256 * 256;
  cudaStream_t cs = ctx->stream, hs = ctx->h2d_stream;
  RQCD_CU(cudaEventRecord(ctx->ev_start, cs));  // copies must not overtake earlier work queued on the compute stream
  RQCD_CU(cudaStreamWaitEvent(hs, ctx->ev_start, 0));
  RQCD_CU(cudaMemsetAsync(ctx->d_stream, 0, 2 * sizeof(unsigned long long), hs));
  auto copy_chunk = [&](int ci, long long c0, long long c1) -> cudaError_t {
    const long long w = c1 - c0;
    for (int b = 0; b < cc.nblocks && w > 0; b++) {
      const RowBlock& rb = cc.blocks[b];
      if (rb.count >= 0) continue;  // shared rows went ahead of the first chunk
      cudaError_t e = cudaMemcpy2DAsync(rb.dst + c0, (size_t)n * 8, rb.src + c0, (size_t)rb.src_stride * 8,
                                        (size_t)w * 8, rb.rows, cudaMemcpyHostToDevice, hs);
      if (e != cudaSuccess) return e;
    }
    ctx->h_stream[ci] = (unsigned long long)c1;
    return cudaMemcpyAsync(ctx->d_stream, ctx->h_stream + ci, sizeof(unsigned long long), cudaMemcpyHostToDevice, hs);
  };
  // Every chunk copy and progress update is enqueued BEFORE the kernel is launched: the copies are asynchronous and
  // ordered on the copy stream, the compute stream only waits for the first chunk. After its launch the kernel never
  // depends on further host progress, so a blocking or serialised launch (CUDA_LAUNCH_BLOCKING, a profiler's kernel
  // replay, a descheduled host thread) cannot turn a valid call into a time-out.
  for (int b = 0; b < cc.nblocks; b++) {
    const RowBlock& rb = cc.blocks[b];
    if (rb.count > 0)
      RQCD_CU(cudaMemcpy2DAsync(rb.dst, (size_t)n * 8, rb.src, (size_t)rb.src_stride * 8, (size_t)rb.count * 8, rb.rows,
                                cudaMemcpyHostToDevice, hs));
  }
  const long long first_end = chunk < n ? chunk : n;
  RQCD_CU(copy_chunk(0, 0, first_end));
  RQCD_CU(cudaEventRecord(ctx->ev_in, hs));
  int ci = 1;
  for (long long c0 = first_end; c0 < n; c0 += chunk, ci++) RQCD_CU(copy_chunk(ci, c0, c0 + chunk < n ? c0 + chunk : n));
  RQCD_CU(cudaStreamWaitEvent(cs, ctx->ev_in, 0));
  cc.p.progress = ctx->d_stream;
  cc.p.stream_error = ctx->d_stream + 1;
  int st = run_check(ctx, cc);
  cc.p.progress = nullptr;
  cc.p.stream_error = nullptr;
  if (st != RQCD_OK) return st;
  st = outputs_d2h(ctx, out, n, m, o);
  if (st != RQCD_OK) return st;
  RQCD_CU(cudaMemcpyAsync(ctx->h_stream + kMaxChunks, ctx->d_stream + 1, sizeof(unsigned long long),
                          cudaMemcpyDeviceToHost, cs));
  RQCD_CU(cudaStreamSynchronize(cs));
  RQCD_CU(cudaStreamSynchronize(hs));
  if (ctx->h_stream[kMaxChunks] != 0) {
    snprintf(ctx->last_error, sizeof ctx->last_error, "streaming input: a chunk did not arrive within the time-out");
    return RQCD_ERR_CUDA;
  }
  return RQCD_OK;
}

// RQCD_F_FRAGILE: after the base launch, k_fragile walks every pair's recursion in lockstep with eight copies whose
// transcendental outputs (the cubic's three cosines, Cardano's cube root) are moved by +-eta, once per probe magnitude
// (RQCD_FRAGILE_LOG2, default 2^-50: ~8 ulp, where two conforming math libraries differ by one or two), and marks the
// pairs with a sign decision that is not pinned (see k_fragile).
int run_fragile_probes(rqcd_ctx* ctx, CheckCall& cc, const OutStage& o, long long n, uint8_t* d_fragile) {
  RQCD_CU(cudaMemsetAsync(d_fragile, 0, (size_t)n, ctx->stream));
  CheckParams q = cc.p;
  q.progress = nullptr;  // the input is all there once the base launch has run
  q.stream_error = nullptr;
  q.list = nullptr;  // one thread per pair of the whole range; pairs of candidates the input filter rejected are skipped
  q.n_list = nullptr;
  if (cc.filter_mask) q.skip = cc.filter_mask;
  const bool moving = ctx->any_moving;
  for (int k = 0; k < ctx->n_probe_mag; k++) {
    RQCD_CU(launch_fragile(q, cc.mode, moving && !cc.pairwise, cc.pairwise, std::ldexp(1.0, ctx->probe_log2[k]),
                           ctx->probe_margin, ctx->probe_floor, cc.pairwise ? nullptr : o.first, d_fragile, cc.pairwise ? nullptr : o.matrix, ctx->d_why,
                           ctx->stream));
    ctx->launches++;
  }
  return RQCD_OK;
}

// Shared tail of the three check entry points.
int finish_check(rqcd_ctx* ctx, CheckCall& cc, const rqcd_out* out, uint32_t flags, long long n, int m,
                 bool want_cost_coeffs) {
  OutStage o;
  const bool dev = flags & RQCD_F_DEVICE_PTRS;
  const bool fragile = (flags & RQCD_F_FRAGILE) != 0;
  if (dev) {
    o.verdict = out->verdict;
    o.first = out->first_obstacle;
    o.matrix = out->verdict_matrix;
    if (want_cost_coeffs) {
      o.cost = out->cost;
      o.coeffs = out->coeffs;
      o.coeff_stride = out->coeff_stride;
    }
  } else {
    // Direct result stores only where the kernel writes a whole batch of 32 candidates at once (k_pool: 32 / 128 / 256-byte
    // coalesced stores); k_check's and the input filter's lanes finish one by one, which would be single small PCIe writes.
    // The fragile probes read the verdicts and first hits back on the device: staged then.
    const bool batch_stores = pool_threads_for(ctx, cc.pairwise) > 0 && !(ctx->filter_on && cc.mode == MODE_BC);
    int st = stage_outputs(ctx, out, n, m, want_cost_coeffs, o, /*allow_direct=*/!fragile && batch_stores);
    if (st != RQCD_OK) return st;
  }
  uint8_t* d_fragile = nullptr;
  if (fragile) {
    // the probes are compared by verdict and first hit: both exist on the device even if the caller did not ask for them
    if (!o.verdict) {
      RQCD_CU(ensure(ctx->s_verdict, (size_t)(n > 0 ? n : 1)));
      o.verdict = (uint8_t*)ctx->s_verdict.p;
    }
    if (!o.first && !cc.pairwise) {
      RQCD_CU(ensure(ctx->s_first, (size_t)(n > 0 ? n : 1) * 4));
      o.first = (int*)ctx->s_first.p;
    }
    d_fragile = out->fragile;
    if (!dev) {
      RQCD_CU(ensure(ctx->s_fragile, (size_t)(n > 0 ? n : 1)));
      d_fragile = (uint8_t*)ctx->s_fragile.p;
    }
  }
  fill_params_out(cc.p, o);
  cc.p.early_exit = (flags & RQCD_F_EARLY_EXIT) ? 1 : 0;
  cc.fast_fma = (flags & RQCD_F_FAST_FMA) != 0;
  int st;
  const bool filter = ctx->filter_on && cc.mode == MODE_BC && n > 0;
  if (filter) {
    uint8_t* d_feas = dev ? out->input_feasibility : nullptr;
    if (!d_feas) {
      RQCD_CU(ensure(ctx->s_feas, (size_t)n));
      d_feas = (uint8_t*)ctx->s_feas.p;
    }
    // the filter reads every candidate before the check starts: the input goes over in full first (no streaming)
    if (!dev) RQCD_CU(copy_blocks(cc, n, ctx->stream));
    st = run_input_filter(ctx, cc, o, n, d_feas);
    if (st == RQCD_OK) st = run_check(ctx, cc);
    if (st == RQCD_OK && !dev) {
      st = outputs_d2h(ctx, out, n, m, o);
      if (st == RQCD_OK && out->input_feasibility)
        RQCD_CU(cudaMemcpyAsync(out->input_feasibility, d_feas, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    }
  } else {
    st = dev ? run_check(ctx, cc) : run_host_pipeline(ctx, cc, out, n, m, o);
  }
  if (st != RQCD_OK) return st;
  int sum_st = RQCD_OK;
  if (out->summary) {
    sum_st = fetch_summary(ctx, out->summary);
    if (sum_st != RQCD_OK && sum_st != RQCD_ERR_DEPTH_CAP) return sum_st;
  }
  if (fragile && n > 0) {
    st = run_fragile_probes(ctx, cc, o, n, d_fragile);
    if (st != RQCD_OK) return st;
    if (!dev) {
      RQCD_CU(cudaMemcpyAsync(out->fragile, d_fragile, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
      if (out->verdict_matrix && m > 0)  // again: bit 7 now marks the fragile pairs
        RQCD_CU(cudaMemcpyAsync(out->verdict_matrix, o.matrix, (size_t)n * (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));
    }
  }
  cc.p.skip = nullptr;
  cc.p.list = nullptr;
  cc.p.n_list = nullptr;
  cc.filter_mask = nullptr;
  if (!dev) RQCD_CU(cudaStreamSynchronize(ctx->stream));
  return sum_st;
}

}  // namespace

extern "C" {

int rqcd_create(rqcd_ctx** out, int device) {
  if (!out) return RQCD_ERR_INVALID_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    return RQCD_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= ndev) return RQCD_ERR_INVALID_ARG;
  rqcd_ctx* ctx = new (std::nothrow) rqcd_ctx();
  if (!ctx) return RQCD_ERR_NOMEM;
  ctx->device = device;
  ctx->stride = OF_STATIC_FIELDS;
  if (cudaGetDeviceProperties(&ctx->prop, device) != cudaSuccess || ctx->prop.major < 10) {
    cudaGetLastError();
    delete ctx;
    return RQCD_ERR_NO_DEVICE;  // the kernels are built for sm_100a only
  }
  DeviceGuard g(device);
  cudaError_t e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_in, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_start, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaMalloc(&ctx->d_counter, sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMalloc(&ctx->d_why, 32 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemset(ctx->d_why, 0, 32 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMalloc(&ctx->d_summary, sizeof(DevSummary));
  if (e == cudaSuccess) e = cudaMallocHost(&ctx->h_summary, sizeof(DevSummary));
  if (e == cudaSuccess) e = cudaMalloc(&ctx->d_stream, 2 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMallocHost(&ctx->h_stream, (kMaxChunks + 1) * sizeof(unsigned long long));
  if (e == cudaSuccess) e = init_device_constants(ctx->own_stream);
  if (e == cudaSuccess) e = init_pool_constants(ctx->own_stream);
  if (e == cudaSuccess) e = rqcd_fma::init_device_constants(ctx->own_stream);
  if (e == cudaSuccess) e = rqcd_fma::init_pool_constants(ctx->own_stream);
  if (e != cudaSuccess) {
    rqcd_destroy(ctx);
    return RQCD_ERR_CUDA;
  }
  ctx->stream = ctx->own_stream;
  const char* hc = getenv("RQCD_HOST_CHUNK");
  if (hc && atoll(hc) > 0) ctx->chunk_override = atoll(hc);
  const char* rm = getenv("RQCD_REFILL_MIN");
  if (rm && atoi(rm) >= 1 && atoi(rm) <= 32) ctx->refill_min = atoi(rm);
  const char* po = getenv("RQCD_POOL");
  if (po && (po[0] == '0' || po[0] == '1' || po[0] == '2') && po[1] == 0) ctx->pool = po[0] - '0';  // 2: k_pool for short tables too
  const char* th = getenv("RQCD_TH_SOLVE");
  if (th && atoi(th) >= 1 && atoi(th) <= 32) ctx->th_solve = atoi(th);
  th = getenv("RQCD_TH_PLANE");
  if (th && atoi(th) >= 1 && atoi(th) <= 32) ctx->th_plane = atoi(th);
  const char* to = getenv("RQCD_STREAM_TIMEOUT_CYCLES");
  if (to && atoll(to) > 0) ctx->stream_timeout = atoll(to);
  const char* fl = getenv("RQCD_FRAGILE_LOG2");  // magnitudes of the RQCD_F_FRAGILE probes, e.g. "-50,-47,-44"
  if (fl && *fl) {
    ctx->n_probe_mag = 0;
    for (const char* q = fl; *q && ctx->n_probe_mag < 8;) {
      char* end = nullptr;
      const long v = strtol(q, &end, 10);
      if (end == q) break;
      if (v <= -30 && v >= -60) ctx->probe_log2[ctx->n_probe_mag++] = (int)v;
      q = (*end == ',') ? end + 1 : end;
    }
    if (ctx->n_probe_mag == 0) ctx->n_probe_mag = 1, ctx->probe_log2[0] = -50;
  }
  const char* pt = getenv("RQCD_POOL_THREADS");
  if (pt && (atoi(pt) == 256 || atoi(pt) == 384 || atoi(pt) == 512)) ctx->pool_threads_override = atoi(pt);
  const char* dz = getenv("RQCD_DIRECT_OUT");
  if (dz && (dz[0] == '0' || dz[0] == '1') && dz[1] == 0) ctx->direct_out = dz[0] == '1';
  const char* fm = getenv("RQCD_FRAGILE_MARGIN");
  if (fm && atof(fm) >= 1.0) ctx->probe_margin = atof(fm);
  const char* ff = getenv("RQCD_FRAGILE_FLOOR");
  if (ff && atof(ff) >= 0.0) ctx->probe_floor = atof(ff);
  const char* ls = getenv("RQCD_LOCKSTEP");
  if (ls && (ls[0] == '0' || ls[0] == '1') && ls[1] == 0) ctx->lockstep = ls[0] - '0';
  *out = ctx;
  return RQCD_OK;
}

void rqcd_destroy(rqcd_ctx* ctx) {
  if (!ctx) return;
  DeviceGuard g(ctx->device);
  if (ctx->own_stream) cudaStreamSynchronize(ctx->own_stream);
  rqcd_comm_finalize(ctx);
  if (ctx->d_merged) cudaFree(ctx->d_merged);
  release(ctx->obs_blob);
  DevBuf* bufs[] = {&ctx->s_in,     &ctx->s_aux,   &ctx->s_verdict, &ctx->s_first,  &ctx->s_matrix, &ctx->s_cost,
                    &ctx->s_coeffs, &ctx->s_roots, &ctx->s_count,   &ctx->s_planes, &ctx->s_peaks,  &ctx->s_group,
                    &ctx->s_fragile, &ctx->s_feas, &ctx->s_list};
  for (DevBuf* b : bufs) release(*b);
  if (ctx->d_counter) cudaFree(ctx->d_counter);
  if (ctx->d_why) cudaFree(ctx->d_why);
  if (ctx->d_partials) cudaFree(ctx->d_partials);
  if (ctx->d_summary) cudaFree(ctx->d_summary);
  if (ctx->h_summary) cudaFreeHost(ctx->h_summary);
  if (ctx->d_stream) cudaFree(ctx->d_stream);
  if (ctx->h_stream) cudaFreeHost(ctx->h_stream);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  if (ctx->h2d_stream) cudaStreamDestroy(ctx->h2d_stream);
  for (cudaEvent_t ev : {ctx->ev_in, ctx->ev_start})
    if (ev) cudaEventDestroy(ev);
  delete ctx;
}

const char* rqcd_strerror(int status) {
  switch (status) {
    case RQCD_OK: return "ok";
    case RQCD_ERR_INVALID_ARG: return "invalid argument";
    case RQCD_ERR_CUDA: return "CUDA runtime error (see rqcd_last_error)";
    case RQCD_ERR_NO_DEVICE: return "no usable sm_100 CUDA device (this library has no CPU fallback)";
    case RQCD_ERR_DEPTH_CAP: return "interval-stack cap hit: (t_end - t_start)/min_time_section is too large";
    case RQCD_ERR_NOMEM: return "out of memory";
    case RQCD_ERR_UNSUPPORTED: return "unsupported obstacle type or obstacle set too large for shared memory";
    default: return "unknown status";
  }
}

const char* rqcd_last_error(const rqcd_ctx* ctx) { return ctx ? ctx->last_error : ""; }

int rqcd_device_info(const rqcd_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, char* name, size_t name_len) {
  if (!ctx) return RQCD_ERR_INVALID_ARG;
  if (sm_count) *sm_count = ctx->prop.multiProcessorCount;
  if (cc_major) *cc_major = ctx->prop.major;
  if (cc_minor) *cc_minor = ctx->prop.minor;
  if (name && name_len > 0) {
    strncpy(name, ctx->prop.name, name_len - 1);
    name[name_len - 1] = 0;
  }
  return RQCD_OK;
}

void* rqcd_get_stream(const rqcd_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int rqcd_set_stream(rqcd_ctx* ctx, void* cuda_stream) {
  if (!ctx) return RQCD_ERR_INVALID_ARG;
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return RQCD_OK;
}

int rqcd_synchronize(rqcd_ctx* ctx) {
  if (!ctx) return RQCD_ERR_INVALID_ARG;
  DeviceGuard g(ctx->device);
  RQCD_CU(cudaStreamSynchronize(ctx->stream));
  return RQCD_OK;
}

int rqcd_read_summary(rqcd_ctx* ctx, rqcd_summary* out) {
  if (!ctx || !out || !ctx->summary_valid) return RQCD_ERR_INVALID_ARG;
  DeviceGuard g(ctx->device);
  return fetch_summary(ctx, out);
}

uint64_t rqcd_launch_count(const rqcd_ctx* ctx) { return ctx ? ctx->launches : 0; }

int rqcd_set_input_filter(rqcd_ctx* ctx, const rqcd_input_limits* limits) {
  if (!ctx) return RQCD_ERR_INVALID_ARG;
  if (!limits) {
    ctx->filter_on = false;
    return RQCD_OK;
  }
  if (bad_min_t(limits->min_time_section)) return RQCD_ERR_INVALID_ARG;
  ctx->filter = *limits;
  ctx->filter_on = true;
  return RQCD_OK;
}

const char* rqcd_last_kernel_name(const rqcd_ctx* ctx) { return ctx ? ctx->last_kernel : ""; }

int rqcd_fragile_reasons(rqcd_ctx* ctx, uint64_t counts[32]) {
  if (!ctx || !counts) return RQCD_ERR_INVALID_ARG;
  DeviceGuard g(ctx->device);
  RQCD_CU(cudaStreamSynchronize(ctx->stream));
  RQCD_CU(cudaMemcpy(counts, ctx->d_why, 32 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return RQCD_OK;
}

int rqcd_set_obstacles(rqcd_ctx* ctx, const rqcd_obstacle* obstacles, int32_t m) {
  if (!ctx || m < 0 || (m > 0 && !obstacles)) return RQCD_ERR_INVALID_ARG;
  DeviceGuard g(ctx->device);
  bool any_moving = false;
  for (int k = 0; k < m; k++) {
    int st = validate_obstacle(obstacles[k]);
    if (st != RQCD_OK) return st;
    any_moving = any_moving || obstacles[k].moving;
  }
  const int stride = (any_moving ? OF_MOVING_FIELDS : OF_STATIC_FIELDS) | 1;  // odd: bank spread, see rqcd_device.cuh
  const int mp = (m + 3) / 4 * 4;
  if (check_smem_bytes(mp, stride, any_moving) > kMaxSmem && pool_threads(mp, stride, kMaxSmem, kMaxSmem) == 0)
    return RQCD_ERR_UNSUPPORTED;
  const size_t bytes = (size_t)stride * mp * 8 + (size_t)mp * 4;
  std::vector<unsigned char> blob(bytes > 0 ? bytes : 16, 0);
  double* tab = reinterpret_cast<double*>(blob.data());
  int* fl = reinterpret_cast<int*>(blob.data() + (size_t)stride * mp * 8);
  for (int k = 0; k < m; k++) {
    const rqcd_obstacle& o = obstacles[k];
    auto F = [&](int f) -> double& { return tab[(size_t)k * stride + f]; };
    F(OF_CX) = o.center[0];
    F(OF_CY) = o.center[1];
    F(OF_CZ) = o.center[2];
    int flags = 0;
    if (o.type == RQCD_SPHERE) {
      // stored as a degenerate prism (identity rotations, zero half sides) + radius: see frame()
      F(OF_RADIUS) = o.radius;
      F(OF_R + 0) = F(OF_R + 4) = F(OF_R + 8) = 1.0;
      F(OF_RI + 0) = F(OF_RI + 4) = F(OF_RI + 8) = 1.0;
    } else {
      flags |= OBS_PRISM;
      for (int i = 0; i < 3; i++) F(OF_H0 + i) = o.side[i] / 2;  // RectPrism.hpp:72-75, :97-99
      double R[9], Ri[9];
      const double qi[4] = {o.quat[0], -o.quat[1], -o.quat[2], -o.quat[3]};  // Rotation::Inverse, Rotation.hpp:63-65
      rotation_matrix(o.quat, R);
      rotation_matrix(qi, Ri);
      for (int i = 0; i < 9; i++) {
        F(OF_R + i) = R[i];
        F(OF_RI + i) = Ri[i];
      }
    }
    {
      // ball around the centre that holds the obstacle, padded (far_from_obstacle in rqcd_device.cuh): the sphere's
      // radius or the norm of the prism's half sides; none (+inf) unless the quaternion is a unit quaternion, i.e. the
      // matrices above are rotations
      double rb = o.radius;
      if (o.type == RQCD_RECT_PRISM) {
        const double h0 = o.side[0] / 2, h1 = o.side[1] / 2, h2 = o.side[2] / 2;
        rb = std::sqrt(h0 * h0 + h1 * h1 + h2 * h2);
        const double q2 = o.quat[0] * o.quat[0] + o.quat[1] * o.quat[1] + o.quat[2] * o.quat[2] + o.quat[3] * o.quat[3];
        if (!(std::fabs(q2 - 1.0) <= 1e-12)) rb = std::numeric_limits<double>::infinity();
      }
      const double cn = std::fabs(o.center[0]) + std::fabs(o.center[1]) + std::fabs(o.center[2]) + 2 * rb;
      F(OF_RBOUND) = rb * (1.0 + 2e-7) + 1e-10 * cn;  // 2e-7: a sphere's plane point is centre + r n with |n| = 1 +- 6e-8
      if (o.moving) F(OF_RBOUND) = std::numeric_limits<double>::infinity();  // the cull is for static obstacles
    }
    if (o.moving) {
      flags |= OBS_MOVING;
      for (int i = 0; i < 18; i++) F(OF_MOTION + i) = o.motion[i];
      F(OF_MT0) = o.motion_t_start;
      F(OF_MT1) = o.motion_t_end;
    }
    fl[k] = flags;
  }
  RQCD_CU(ensure(ctx->obs_blob, blob.size()));
  // synchronous on purpose: the blob is a stack-lifetime host vector, and a previous launch may still read the old set
  RQCD_CU(cudaStreamSynchronize(ctx->stream));
  RQCD_CU(cudaMemcpy(ctx->obs_blob.p, blob.data(), blob.size(), cudaMemcpyHostToDevice));
  ctx->m = m;
  ctx->mp = mp;
  ctx->stride = stride;
  ctx->any_moving = any_moving;
  return RQCD_OK;
}

int32_t rqcd_num_obstacles(const rqcd_ctx* ctx) { return ctx ? ctx->m : -1; }

int rqcd_generate(rqcd_ctx* ctx, const rqcd_bc_soa* in, int64_t n, uint32_t flags, rqcd_out* out) {
  if (!ctx || !out || bad_bc(in, n)) return RQCD_ERR_INVALID_ARG;
  if (out->coeffs && out->coeff_stride < n) return RQCD_ERR_INVALID_ARG;
  if (out->params && out->param_stride < n) return RQCD_ERR_INVALID_ARG;
  if (n == 0) return RQCD_OK;
  DeviceGuard g(ctx->device);
  CheckParams p = {};
  p.n = n;
  if (flags & RQCD_F_DEVICE_PTRS) {
    bind_bc(p, in);
    RQCD_CU(launch_generate(p, out->cost, out->coeffs, out->coeff_stride, out->params, out->param_stride, ctx->stream));
    ctx->launches++;
    return RQCD_OK;
  }
  RQCD_CU(ensure(ctx->s_in, (size_t)n * 8 * RQCD_BC_ROWS));
  BlockList bl;
  stage_bc(p, bl, in, n, ctx->s_in.p);
  RQCD_CU(copy_blocks(bl, n, ctx->stream));
  OutStage o;
  if (out->cost) {
    RQCD_CU(ensure(ctx->s_cost, (size_t)n * 8));
    o.cost = (double*)ctx->s_cost.p;
  }
  if (out->coeffs) {
    RQCD_CU(ensure(ctx->s_coeffs, (size_t)n * 8 * RQCD_COEFF_ROWS));
    o.coeffs = (double*)ctx->s_coeffs.p;
    o.coeff_stride = n;
  }
  double* d_params = nullptr;
  if (out->params) {
    RQCD_CU(ensure(ctx->s_roots, (size_t)n * 8 * RQCD_PARAM_ROWS));
    d_params = (double*)ctx->s_roots.p;
  }
  RQCD_CU(launch_generate(p, o.cost, o.coeffs, o.coeff_stride, d_params, n, ctx->stream));
  ctx->launches++;
  if (o.cost) RQCD_CU(cudaMemcpyAsync(out->cost, o.cost, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (o.coeffs) RQCD_CU(rows_d2h(out->coeffs, out->coeff_stride, o.coeffs, n, RQCD_COEFF_ROWS, ctx->stream));
  if (d_params) RQCD_CU(rows_d2h(out->params, out->param_stride, d_params, n, RQCD_PARAM_ROWS, ctx->stream));
  RQCD_CU(cudaStreamSynchronize(ctx->stream));
  return RQCD_OK;
}

int rqcd_generate_check(rqcd_ctx* ctx, const rqcd_bc_soa* in, int64_t n, double min_time_section, uint32_t flags,
                        int64_t index_base, rqcd_out* out) {
  if (!ctx || bad_bc(in, n) || bad_min_t(min_time_section)) return RQCD_ERR_INVALID_ARG;
  int st = validate_out(out, flags, n);
  if (st != RQCD_OK) return st;
  DeviceGuard g(ctx->device);
  CheckCall cc;
  cc.mode = MODE_BC;
  cc.p.n = n;
  cc.p.index_base = index_base;
  cc.p.min_t = min_time_section;
  if (flags & RQCD_F_DEVICE_PTRS) {
    bind_bc(cc.p, in);
  } else {
    RQCD_CU(ensure(ctx->s_in, (size_t)n * 8 * RQCD_BC_ROWS));
    stage_bc(cc.p, cc, in, n, ctx->s_in.p);
  }
  return finish_check(ctx, cc, out, flags, n, ctx->m, true);
}

int rqcd_check(rqcd_ctx* ctx, const rqcd_coeff_soa* in, const double* cost, int64_t n, double min_time_section,
               uint32_t flags, int64_t index_base, rqcd_out* out) {
  if (!ctx || !in || n < 0 || (n > 0 && (!in->coeffs || !in->t_end)) || in->stride < n || bad_min_t(min_time_section))
    return RQCD_ERR_INVALID_ARG;
  int st = validate_out(out, flags, n);
  if (st != RQCD_OK) return st;
  DeviceGuard g(ctx->device);
  CheckCall cc;
  cc.mode = MODE_COEFF;
  cc.p.n = n;
  cc.p.index_base = index_base;
  cc.p.min_t = min_time_section;
  if (flags & RQCD_F_DEVICE_PTRS) {
    cc.p.in = in->coeffs;
    cc.p.in_stride = in->stride;
    cc.p.t_start = in->t_start;
    cc.p.t_end = in->t_end;
    cc.p.cost_in = cost;
  } else {
    RQCD_CU(ensure(ctx->s_in, (size_t)n * 8 * RQCD_COEFF_ROWS));
    RQCD_CU(ensure(ctx->s_aux, (size_t)n * 8 * 3));
    double* aux = (double*)ctx->s_aux.p;
    cc.add(ctx->s_in.p, in->coeffs, in->stride, RQCD_COEFF_ROWS);
    cc.add(aux, in->t_end, n, 1);
    if (in->t_start) cc.add(aux + n, in->t_start, n, 1);
    if (cost) cc.add(aux + 2 * n, cost, n, 1);
    cc.p.in = (const double*)ctx->s_in.p;
    cc.p.in_stride = n;
    cc.p.t_end = aux;
    cc.p.t_start = in->t_start ? aux + n : nullptr;
    cc.p.cost_in = cost ? aux + 2 * n : nullptr;
  }
  // Trajectory's constructor asserts startTime <= endTime (Trajectory.hpp:68); windows that violate it are
  // checked like any other input here (the section loop simply sees tf - ts < minT).
  return finish_check(ctx, cc, out, flags, n, ctx->m, false);
}

int rqcd_generate_check_pairwise(rqcd_ctx* ctx, const rqcd_bc_soa* in, const double* spheres, int64_t sphere_stride,
                                 int64_t n, double min_time_section, uint32_t flags, int64_t index_base, rqcd_out* out) {
  if (!ctx || bad_bc(in, n) || (n > 0 && !spheres) || sphere_stride < n || bad_min_t(min_time_section))
    return RQCD_ERR_INVALID_ARG;
  int st = validate_out(out, flags, n);
  if (st != RQCD_OK) return st;
  if (out->verdict_matrix || out->first_obstacle) return RQCD_ERR_INVALID_ARG;
  DeviceGuard g(ctx->device);
  CheckCall cc;
  cc.mode = MODE_BC;
  cc.pairwise = true;
  cc.p.n = n;
  cc.p.index_base = index_base;
  cc.p.min_t = min_time_section;
  if (flags & RQCD_F_DEVICE_PTRS) {
    bind_bc(cc.p, in);
    cc.p.spheres = spheres;
    cc.p.sphere_stride = sphere_stride;
  } else {
    RQCD_CU(ensure(ctx->s_in, (size_t)n * 8 * RQCD_BC_ROWS));
    RQCD_CU(ensure(ctx->s_aux, (size_t)n * 8 * 4));
    stage_bc(cc.p, cc, in, n, ctx->s_in.p);
    cc.add(ctx->s_aux.p, spheres, sphere_stride, 4);
    cc.p.spheres = (const double*)ctx->s_aux.p;
    cc.p.sphere_stride = n;
  }
  return finish_check(ctx, cc, out, flags & ~RQCD_F_EARLY_EXIT, n, 1, true);
}

static int planes_common(rqcd_ctx* ctx, int mode, CheckParams& p, const BlockList& bl, int64_t n,
                         const double* planes, int32_t np, uint32_t flags, uint8_t* verdict, uint8_t* matrix) {
  if (np < 0 || (np > 0 && !planes)) return RQCD_ERR_INVALID_ARG;
  // Boundary.normal = normal.GetUnitVector() (CollisionChecker.cpp:29): float-narrowed norm, Vec3.hpp:117-120
  std::vector<double> prep((size_t)(np > 0 ? np : 1) * 6);
  for (int k = 0; k < np; k++) {
    const double* q = planes + 6 * k;
    for (int i = 0; i < 6; i++)
      if (!std::isfinite(q[i])) return RQCD_ERR_INVALID_ARG;
    const double nn = (double)(float)std::sqrt(q[3] * q[3] + q[4] * q[4] + q[5] * q[5]);
    for (int i = 0; i < 3; i++) {
      prep[6 * k + i] = q[i];
      prep[6 * k + 3 + i] = q[3 + i] / nn;
    }
  }
  RQCD_CU(ensure(ctx->s_planes, prep.size() * 8));
  RQCD_CU(cudaStreamSynchronize(ctx->stream));
  RQCD_CU(cudaMemcpy(ctx->s_planes.p, prep.data(), prep.size() * 8, cudaMemcpyHostToDevice));
  if (n == 0) return RQCD_OK;
  const bool dev = flags & RQCD_F_DEVICE_PTRS;
  uint8_t *dv = verdict, *dm = matrix;
  if (!dev) {
    RQCD_CU(copy_blocks(bl, n, ctx->stream));
    if (verdict) {
      RQCD_CU(ensure(ctx->s_verdict, (size_t)n));
      dv = (uint8_t*)ctx->s_verdict.p;
    }
    if (matrix) {
      RQCD_CU(ensure(ctx->s_matrix, (size_t)n * (size_t)(np > 0 ? np : 1)));
      dm = (uint8_t*)ctx->s_matrix.p;
    }
  }
  p.n = n;
  RQCD_CU(launch_planes(p, mode, (const double*)ctx->s_planes.p, np, dv, dm, ctx->stream));
  ctx->launches++;
  if (!dev) {
    if (verdict) RQCD_CU(cudaMemcpyAsync(verdict, dv, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (matrix && np > 0) RQCD_CU(cudaMemcpyAsync(matrix, dm, (size_t)n * np, cudaMemcpyDeviceToHost, ctx->stream));
    RQCD_CU(cudaStreamSynchronize(ctx->stream));
  }
  return RQCD_OK;
}

int rqcd_check_planes(rqcd_ctx* ctx, const rqcd_coeff_soa* in, int64_t n, const double* planes, int32_t np,
                      uint32_t flags, uint8_t* verdict, uint8_t* matrix) {
  if (!ctx || !in || n < 0 || (n > 0 && (!in->coeffs || !in->t_end)) || in->stride < n) return RQCD_ERR_INVALID_ARG;
  DeviceGuard g(ctx->device);
  CheckParams p = {};
  BlockList bl;
  if (flags & RQCD_F_DEVICE_PTRS) {
    p.in = in->coeffs;
    p.in_stride = in->stride;
    p.t_start = in->t_start;
    p.t_end = in->t_end;
  } else {
    RQCD_CU(ensure(ctx->s_in, (size_t)n * 8 * RQCD_COEFF_ROWS));
    RQCD_CU(ensure(ctx->s_aux, (size_t)n * 8 * 2));
    double* aux = (double*)ctx->s_aux.p;
    bl.add(ctx->s_in.p, in->coeffs, in->stride, RQCD_COEFF_ROWS);
    bl.add(aux, in->t_end, n, 1);
    if (in->t_start) bl.add(aux + n, in->t_start, n, 1);
    p.in = (const double*)ctx->s_in.p;
    p.in_stride = n;
    p.t_end = aux;
    p.t_start = in->t_start ? aux + n : nullptr;
  }
  return planes_common(ctx, MODE_COEFF, p, bl, n, planes, np, flags, verdict, matrix);
}

int rqcd_generate_check_planes(rqcd_ctx* ctx, const rqcd_bc_soa* in, int64_t n, const double* planes, int32_t np,
                               uint32_t flags, uint8_t* verdict, uint8_t* matrix) {
  if (!ctx || bad_bc(in, n)) return RQCD_ERR_INVALID_ARG;
  DeviceGuard g(ctx->device);
  CheckParams p = {};
  BlockList bl;
  if (flags & RQCD_F_DEVICE_PTRS) {
    bind_bc(p, in);
  } else {
    RQCD_CU(ensure(ctx->s_in, (size_t)n * 8 * RQCD_BC_ROWS));
    stage_bc(p, bl, in, n, ctx->s_in.p);
  }
  return planes_common(ctx, MODE_BC, p, bl, n, planes, np, flags, verdict, matrix);
}

int rqcd_check_input_feasibility(rqcd_ctx* ctx, const rqcd_bc_soa* in, int64_t n, const double* gravity, double fmin,
                           double fmax, double wmax, double min_time_section, const int64_t* group, uint32_t flags,
                           uint8_t* result) {
  if (!ctx || !gravity || bad_bc(in, n) || (n > 0 && !result) || bad_min_t(min_time_section)) return RQCD_ERR_INVALID_ARG;
  if (n == 0) return RQCD_OK;
  DeviceGuard g(ctx->device);
  const bool dev = flags & RQCD_F_DEVICE_PTRS;
  CheckParams p = {};
  p.n = n;
  const long long* dgroup = (const long long*)group;
  uint8_t* dres = result;
  if (dev) {
    bind_bc(p, in);
  } else {
    RQCD_CU(ensure(ctx->s_in, (size_t)n * 8 * RQCD_BC_ROWS));
    BlockList bl;
    stage_bc(p, bl, in, n, ctx->s_in.p);
    RQCD_CU(copy_blocks(bl, n, ctx->stream));
    if (group) {
      RQCD_CU(ensure(ctx->s_group, (size_t)n * 8));
      RQCD_CU(cudaMemcpyAsync(ctx->s_group.p, group, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
      dgroup = (const long long*)ctx->s_group.p;
    }
    RQCD_CU(ensure(ctx->s_verdict, (size_t)n));
    dres = (uint8_t*)ctx->s_verdict.p;
  }
  if (group) RQCD_CU(ensure(ctx->s_peaks, (size_t)n * 8 * 6 + (size_t)n));  // + the anchor bytes, see k_feas_group_peaks
  p.work_counter = ctx->d_counter;
  RQCD_CU(launch_feasibility(p, gravity, fmin, fmax, wmax, min_time_section, dgroup, (double*)ctx->s_peaks.p, dres,
                             nullptr, nullptr, nullptr, nullptr, ctx->stream));
  ctx->launches += group ? 2 : 1;
  if (!dev) {
    RQCD_CU(cudaMemcpyAsync(result, dres, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    RQCD_CU(cudaStreamSynchronize(ctx->stream));
  }
  return RQCD_OK;
}

int rqcd_plan_best(rqcd_ctx* ctx, const rqcd_bc_soa* in, int64_t n, const rqcd_plan_params* prm, uint32_t flags,
                   int64_t index_base, uint8_t* stage, rqcd_summary* summary) {
  if (!ctx || !prm || !summary || bad_bc(in, n) || bad_min_t(prm->min_time_section) ||
      prm->n_planes < 0 || (prm->n_planes > 0 && !prm->planes))
    return RQCD_ERR_INVALID_ARG;
  DeviceGuard g(ctx->device);
  const bool dev = flags & RQCD_F_DEVICE_PTRS;
  cudaStream_t s = ctx->stream;
  // inputs on the device
  CheckParams p = {};
  p.n = n;
  if (dev) {
    bind_bc(p, in);
  } else {
    RQCD_CU(ensure(ctx->s_in, (size_t)n * 8 * RQCD_BC_ROWS));
    BlockList bl;
    stage_bc(p, bl, in, n, ctx->s_in.p);
    RQCD_CU(copy_blocks(bl, n, s));
  }
  // scratch: cost, feasibility, plane hits, stage, verdict
  RQCD_CU(ensure(ctx->s_cost, (size_t)(n > 0 ? n : 1) * 8));
  RQCD_CU(ensure(ctx->s_aux, (size_t)(n > 0 ? n : 1) * 4));
  RQCD_CU(ensure(ctx->s_verdict, (size_t)(n > 0 ? n : 1)));
  double* d_cost = (double*)ctx->s_cost.p;
  uint8_t* d_feas = (uint8_t*)ctx->s_aux.p;
  uint8_t* d_plane = d_feas + n;
  uint8_t* d_stage_scratch = d_feas + 2 * n;
  uint8_t* d_stage = (dev && stage) ? stage : d_stage_scratch;
  uint8_t* d_verdict = (uint8_t*)ctx->s_verdict.p;
  // stage 0: Generate + GetCost
  RQCD_CU(launch_generate(p, d_cost, nullptr, 0, nullptr, 0, s));
  ctx->launches += n > 0;
  // stage 1: CheckInputFeasibility
  p.work_counter = ctx->d_counter;
  if (prm->check_input && n > 0) {
    RQCD_CU(launch_feasibility(p, prm->gravity, prm->fmin, prm->fmax, prm->wmax, prm->min_time_section, nullptr, nullptr,
                               d_feas, nullptr, nullptr, nullptr, nullptr, s));
    ctx->launches++;
  }
  // stage 2: plane boundaries
  if (prm->n_planes > 0) {
    std::vector<double> prep((size_t)prm->n_planes * 6);
    for (int k = 0; k < prm->n_planes; k++) {
      const double* q = prm->planes + 6 * k;
      for (int i = 0; i < 6; i++)
        if (!std::isfinite(q[i])) return RQCD_ERR_INVALID_ARG;
      const double nn = (double)(float)std::sqrt(q[3] * q[3] + q[4] * q[4] + q[5] * q[5]);
      for (int i = 0; i < 3; i++) prep[6 * k + i] = q[i], prep[6 * k + 3 + i] = q[3 + i] / nn;
    }
    RQCD_CU(ensure(ctx->s_planes, prep.size() * 8));
    RQCD_CU(cudaStreamSynchronize(s));
    RQCD_CU(cudaMemcpy(ctx->s_planes.p, prep.data(), prep.size() * 8, cudaMemcpyHostToDevice));
    if (n > 0) {
      RQCD_CU(launch_planes(p, MODE_BC, (const double*)ctx->s_planes.p, prm->n_planes, d_plane, nullptr, s));
      ctx->launches++;
    }
  }
  RQCD_CU(launch_plan_mask(n, d_cost, prm->max_cost, prm->check_input ? d_feas : nullptr,
                           prm->n_planes > 0 ? d_plane : nullptr, d_stage, s));
  ctx->launches += n > 0;
  // stage 3: obstacles, only for the survivors; argmin cost among those that stay collision free
  CheckCall cc;
  cc.mode = MODE_BC;
  cc.p = p;
  cc.p.index_base = index_base;
  cc.p.min_t = prm->min_time_section;
  cc.p.skip = d_stage;
  cc.p.verdict = d_verdict;
  cc.p.early_exit = 1;
  if (n > 0) RQCD_CU(cudaMemsetAsync(d_verdict, 0, (size_t)n, s));
  int st = run_check(ctx, cc);
  if (st != RQCD_OK) return st;
  RQCD_CU(launch_plan_stage(n, d_verdict, d_stage, s));
  ctx->launches += n > 0;
  if (stage && !(dev) && n > 0) RQCD_CU(cudaMemcpyAsync(stage, d_stage, (size_t)n, cudaMemcpyDeviceToHost, s));
  return fetch_summary(ctx, summary);
}

static int solve_common(rqcd_ctx* ctx, bool quartic, const double* coef, int64_t n, uint32_t flags, double* roots,
                        int32_t* count) {
  if (!ctx || n < 0 || (n > 0 && (!coef || !roots || !count))) return RQCD_ERR_INVALID_ARG;
  if (n == 0) return RQCD_OK;
  DeviceGuard g(ctx->device);
  const int rows = quartic ? 4 : 3;
  if (flags & RQCD_F_DEVICE_PTRS) {
    RQCD_CU(launch_solve(quartic, coef, n, roots, count, ctx->stream));
    ctx->launches++;
    return RQCD_OK;
  }
  RQCD_CU(ensure(ctx->s_in, (size_t)n * 8 * rows));
  RQCD_CU(ensure(ctx->s_roots, (size_t)n * 8 * rows));
  RQCD_CU(ensure(ctx->s_count, (size_t)n * 4));
  RQCD_CU(cudaMemcpyAsync(ctx->s_in.p, coef, (size_t)n * 8 * rows, cudaMemcpyHostToDevice, ctx->stream));
  RQCD_CU(launch_solve(quartic, (const double*)ctx->s_in.p, n, (double*)ctx->s_roots.p, (int*)ctx->s_count.p, ctx->stream));
  ctx->launches++;
  RQCD_CU(cudaMemcpyAsync(roots, ctx->s_roots.p, (size_t)n * 8 * rows, cudaMemcpyDeviceToHost, ctx->stream));
  RQCD_CU(cudaMemcpyAsync(count, ctx->s_count.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  RQCD_CU(cudaStreamSynchronize(ctx->stream));
  return RQCD_OK;
}

int rqcd_solve_cubic(rqcd_ctx* ctx, const double* coef, int64_t n, uint32_t flags, double* roots, int32_t* count) {
  return solve_common(ctx, false, coef, n, flags, roots, count);
}
int rqcd_solve_quartic(rqcd_ctx* ctx, const double* coef, int64_t n, uint32_t flags, double* roots, int32_t* count) {
  return solve_common(ctx, true, coef, n, flags, roots, count);
}

int rqcd_solve_conditioning(rqcd_ctx* ctx, int32_t degree, const double* coef, int64_t n, uint32_t flags, double* cond,
                            uint8_t* branch_fragile) {
  if (!ctx || n < 0 || (degree != 3 && degree != 4) || (n > 0 && (!coef || !cond || !branch_fragile))) return RQCD_ERR_INVALID_ARG;
  if (n == 0) return RQCD_OK;
  DeviceGuard g(ctx->device);
  const bool quartic = degree == 4;
  const double eta = std::ldexp(1.0, ctx->probe_log2[0]);
  if (flags & RQCD_F_DEVICE_PTRS) {
    RQCD_CU(launch_solve_cond(quartic, coef, n, eta, ctx->probe_margin, ctx->probe_floor, cond, branch_fragile, ctx->stream));
    ctx->launches++;
    return RQCD_OK;
  }
  RQCD_CU(ensure(ctx->s_in, (size_t)n * 8 * degree));
  RQCD_CU(ensure(ctx->s_roots, (size_t)n * 8));
  RQCD_CU(ensure(ctx->s_fragile, (size_t)n));
  RQCD_CU(cudaMemcpyAsync(ctx->s_in.p, coef, (size_t)n * 8 * degree, cudaMemcpyHostToDevice, ctx->stream));
  RQCD_CU(launch_solve_cond(quartic, (const double*)ctx->s_in.p, n, eta, ctx->probe_margin, ctx->probe_floor,
                            (double*)ctx->s_roots.p, (uint8_t*)ctx->s_fragile.p, ctx->stream));
  ctx->launches++;
  RQCD_CU(cudaMemcpyAsync(cond, ctx->s_roots.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  RQCD_CU(cudaMemcpyAsync(branch_fragile, ctx->s_fragile.p, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  RQCD_CU(cudaStreamSynchronize(ctx->stream));
  return RQCD_OK;
}

int rqcd_merge_summaries(const rqcd_summary* shards, int32_t n_shards, rqcd_summary* merged) {
  if (!merged || n_shards < 0 || (n_shards > 0 && !shards)) return RQCD_ERR_INVALID_ARG;
  memset(merged, 0, sizeof *merged);
  merged->best_cost = std::numeric_limits<double>::infinity();
  merged->best_index = -1;
  for (int k = 0; k < n_shards; k++) {
    const rqcd_summary& s = shards[k];
    for (int v = 0; v < 4; v++) {
      merged->traj_counts[v] += s.traj_counts[v];
      merged->pair_counts[v] += s.pair_counts[v];
    }
    merged->pairs_checked += s.pairs_checked;
    if (s.best_index >= 0 && (merged->best_index < 0 || s.best_cost < merged->best_cost ||
                              (s.best_cost == merged->best_cost && s.best_index < merged->best_index))) {
      merged->best_cost = s.best_cost;
      merged->best_index = s.best_index;
    }
  }
  return RQCD_OK;
}

int rqcd_synth_bc(rqcd_ctx* ctx, uint64_t seed, int64_t index_base, int64_t n, double* d_bc, int64_t stride) {
  if (!ctx || n < 0 || (n > 0 && !d_bc) || stride < n) return RQCD_ERR_INVALID_ARG;
  if (n == 0) return RQCD_OK;
  DeviceGuard g(ctx->device);
  RQCD_CU(launch_synth_bc(seed, index_base, n, d_bc, stride, ctx->stream));
  ctx->launches++;
  return RQCD_OK;
}

int rqcd_measure_fp64_peak(rqcd_ctx* ctx, int millis, double* dfma_flops, double* dmul_dadd_flops) {
  if (!ctx || millis <= 0) return RQCD_ERR_INVALID_ARG;
  DeviceGuard g(ctx->device);
  double* sink = nullptr;
  RQCD_CU(cudaMalloc(&sink, 8));
  cudaEvent_t e0, e1;
  RQCD_CU(cudaEventCreate(&e0));
  RQCD_CU(cudaEventCreate(&e1));
  const int blocks = ctx->prop.multiProcessorCount * 8;
  int status = RQCD_OK;
  for (int pass = 0; pass < 2 && status == RQCD_OK; pass++) {
    const bool fma = pass == 0;
    double* dst = fma ? dfma_flops : dmul_dadd_flops;
    int iters = 256;
    float ms = 0;
    for (int attempt = 0; attempt < 3; attempt++) {  // warm-up, calibration, measurement
      cudaError_t e = cudaEventRecord(e0, ctx->stream);
      if (e == cudaSuccess) e = launch_fp64_peak(fma, blocks, iters, sink, ctx->stream);
      if (e == cudaSuccess) e = cudaEventRecord(e1, ctx->stream);
      if (e == cudaSuccess) e = cudaEventSynchronize(e1);
      if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
      if (e != cudaSuccess) {
        status = fail_cuda(ctx, e, "fp64 peak");
        break;
      }
      ctx->launches++;
      if (attempt == 1 && ms > 0) {
        double scale = (double)millis / ms;
        long long it = (long long)(iters * scale);
        if (it < 64) it = 64;
        if (it > (1 << 24)) it = 1 << 24;
        iters = (int)it;
      }
    }
    if (status == RQCD_OK && dst) {
      const double ops = (double)blocks * kPeakThreads * kPeakChains * kPeakInnerOps * (double)iters;
      *dst = ops * (fma ? 2.0 : 1.0) / (ms * 1e-3);
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  return status;
}

}  // extern "C"
