// rqcd_host.cuh -- host-side internals shared by the translation units that implement include/rqcd.h (rqcd_api.cu:
// single-device entry points; rqcd_multi.cu: device groups and the NCCL exchange): the context record, CUDA error
// plumbing, the device guard.
#pragma once
#include <cuda_runtime.h>

#include <cstdio>

#include "../../include/rqcd.h"
#include "rqcd_kernels.cuh"

namespace rqcd {

constexpr size_t kMaxSmem = 227 * 1024;
constexpr int kMaxChunks = 64;

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

}  // namespace rqcd

struct rqcd_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  cudaStream_t h2d_stream = nullptr;                         // host-pointer calls: input streams in beside the kernel
  cudaEvent_t ev_in = nullptr, ev_start = nullptr;
  long long chunk_override = 0;                              // RQCD_HOST_CHUNK: candidates per chunk (0 = automatic)
  cudaDeviceProp prop;
  // obstacle set
  int m = 0, mp = 0, stride = 0;  // record stride in doubles (odd); rqcd_create: OF_STATIC_FIELDS
  bool any_moving = false;
  rqcd::DevBuf obs_blob;
  // launch scratch
  unsigned long long* d_counter = nullptr;
  unsigned long long* d_why = nullptr;  // [32] RQCD_F_FRAGILE: marked pairs by the first decision that was not pinned
  rqcd::BlockPartial* d_partials = nullptr;
  int partial_cap = 0;
  rqcd::DevSummary* d_summary = nullptr;
  rqcd::DevSummary* h_summary = nullptr;  // pinned
  unsigned long long* d_stream = nullptr;  // [0] progress, [1] error flag (streaming host-pointer calls)
  unsigned long long* h_stream = nullptr;  // pinned: [0..kMaxChunks) cumulative counts, [kMaxChunks] error read-back
  bool summary_valid = false;
  // exchange between ranks (rqcd_comm_*, rqcd_multi.cu): one process per GPU, NCCL all-gather of the summary records
  void* comm = nullptr;             // ncclComm_t
  int comm_rank = 0, comm_world = 1;
  rqcd::DevSummary* d_gather = nullptr;  // [comm_world] records, this rank's at [comm_rank]
  rqcd::DevSummary* d_merged = nullptr;  // merge of the table (rqcd_exchange_summaries)
  bool merged_valid = false;
  // staging for host-pointer calls
  rqcd::DevBuf s_in, s_aux, s_verdict, s_first, s_matrix, s_cost, s_coeffs, s_roots, s_count, s_planes, s_peaks, s_group, s_fragile, s_feas, s_list;
  int pool_threads_override = 0;  // RQCD_POOL_THREADS
  bool direct_out = true;  // RQCD_DIRECT_OUT: results of host-pointer calls stored straight into page-locked output buffers
  bool filter_on = false;  // rqcd_set_input_filter
  rqcd_input_limits filter = {};
  int refill_min = 0;  // RQCD_REFILL_MIN override; 0 = chosen from the obstacle count
  int lockstep = -1;   // RQCD_LOCKSTEP override (0/1); -1 = chosen from the obstacle count
  int pool = 1;        // RQCD_POOL: 1 = k_pool (staged sections, certified solver skip) for moving tables and static ones of
                       // 16+ obstacles, 2 = for every table, 0 = never (k_check)
  int th_solve = 32, th_plane = 24;  // RQCD_TH_SOLVE / RQCD_TH_PLANE: see CheckParams
  long long stream_timeout = 1ll << 32;  // RQCD_STREAM_TIMEOUT_CYCLES (~2 s): a streamed chunk that never arrives fails the call
  int n_probe_mag = 1, probe_log2[8] = {-50, 0, 0, 0, 0, 0, 0, 0};  // RQCD_F_FRAGILE probe magnitudes 2^k (RQCD_FRAGILE_LOG2)
  double probe_margin = 4.0, probe_floor = 8.0;  // RQCD_FRAGILE_MARGIN, RQCD_FRAGILE_FLOOR (ulps), see k_fragile
  const char* last_kernel = "";  // name of the check kernel of the most recent launch (rqcd_last_kernel_name)
  unsigned long long launches = 0;
  char last_error[512] = {0};
};

namespace rqcd {

inline int fail_cuda(rqcd_ctx* c, cudaError_t e, const char* what) {
  snprintf(c->last_error, sizeof c->last_error, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return RQCD_ERR_CUDA;
}

#define RQCD_CU(call)                                        \
  do {                                                       \
    cudaError_t e_ = (call);                                 \
    if (e_ != cudaSuccess) return fail_cuda(ctx, e_, #call); \
  } while (0)

struct DeviceGuard {  // the context's device is current for the duration of a call
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

}  // namespace rqcd
