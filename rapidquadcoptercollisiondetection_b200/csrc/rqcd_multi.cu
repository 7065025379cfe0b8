// rqcd_multi.cu -- the multi-GPU part of include/rqcd.h. Host logic only; the two kernels it launches (k_publish_summary,
// k_merge_summaries) live in rqcd_kernels.cu.
//
//   rqcd_group_*   one process, several GPUs: a context per device, candidates sharded by index, every shard's
//                  88-byte summary stored into a slot table on the first device through the peer mapping (NVLink) or
//                  mapped host memory, merged there by one kernel.
//   rqcd_comm_*    one process per GPU: a context-owned NCCL communicator; the exchange is ncclAllGather of the records
//                  on the compute stream + the same merge kernel. NCCL is loaded with dlopen when a communicator is
//                  first asked for.
//
// The reduction replaces what the reference's drivers do on the host after every check (counting verdicts,
// test/ForestPerformanceEval.cpp:216-230) and a planner's "cheapest collision-free candidate".
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "../../include/rqcd.h"
#include "rqcd_host.cuh"
#include "rqcd_kernels.cuh"

using namespace rqcd;

static_assert(RQCD_COMM_ID_BYTES == sizeof(ncclUniqueId), "RQCD_COMM_ID_BYTES mirrors ncclUniqueId");

namespace {

// ---- NCCL, bound at run time -------------------------------------------------------------------------------------
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  char error[256] = {0};
};

NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {getenv("RQCD_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      if (!nm || !*nm) continue;
      api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) {
      snprintf(api.error, sizeof api.error, "cannot load NCCL (libnccl.so.2; set RQCD_NCCL_LIB): %s", dlerror());
      return;
    }
    auto sym = [&](const char* n) { return dlsym(api.handle, n); };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    if (!api.GetUniqueId || !api.CommInitRank || !api.AllGather || !api.CommDestroy || !api.GetErrorString) {
      snprintf(api.error, sizeof api.error, "the NCCL library lacks a required symbol");
      api.handle = nullptr;
    }
  });
  return &api;
}

int fail_nccl(rqcd_ctx* ctx, NcclApi* a, ncclResult_t r, const char* what) {
  snprintf(ctx->last_error, sizeof ctx->last_error, "%s: %s", what, a->GetErrorString ? a->GetErrorString(r) : "NCCL error");
  return RQCD_ERR_CUDA;
}

int depth_cap_status(const rqcd_summary& s) {
  return (s.traj_counts[3] != 0 || s.pair_counts[3] != 0) ? RQCD_ERR_DEPTH_CAP : RQCD_OK;
}

}  // namespace

// ---- device group ----------------------------------------------------------------------------------------------------
struct rqcd_group {
  std::vector<rqcd_ctx*> ctx;
  std::vector<cudaEvent_t> published;  // [k] on ctx[k]'s device: shard k's record is in its slot
  DevSummary* slots = nullptr;         // [size] on the first device (peer-mapped for the others), or mapped host memory
  bool peer_slots = true;
  DevSummary* d_merged = nullptr;      // first device
  DevSummary* h_merged = nullptr;      // pinned
  bool merged_valid = false;
  char last_error[512] = {0};
};

namespace {

int group_fail(rqcd_group* g, int status, const char* what, const rqcd_ctx* from) {
  snprintf(g->last_error, sizeof g->last_error, "%s: %s%s%.400s", what, rqcd_strerror(status), from && from->last_error[0] ? ": " : "",
           from ? from->last_error : "");
  return status;
}

int group_fail_cuda(rqcd_group* g, cudaError_t e, const char* what) {
  snprintf(g->last_error, sizeof g->last_error, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return RQCD_ERR_CUDA;
}

#define RQCD_GCU(call)                                         \
  do {                                                         \
    cudaError_t e_ = (call);                                   \
    if (e_ != cudaSuccess) return group_fail_cuda(g, e_, #call); \
  } while (0)

// after shard k's check launch: its record to its slot, and the event the merge waits for
int publish(rqcd_group* g, int k) {
  rqcd_ctx* c = g->ctx[k];
  DeviceGuard guard(c->device);
  RQCD_GCU(launch_publish_summary(c->d_summary, g->slots + k, c->stream));
  c->launches++;
  RQCD_GCU(cudaEventRecord(g->published[k], c->stream));
  return RQCD_OK;
}

// on the first device: wait for every shard's record, merge
int merge(rqcd_group* g) {
  rqcd_ctx* c0 = g->ctx[0];
  DeviceGuard guard(c0->device);
  for (size_t k = 1; k < g->ctx.size(); k++) RQCD_GCU(cudaStreamWaitEvent(c0->stream, g->published[k], 0));
  RQCD_GCU(launch_merge_summaries(g->slots, (int)g->ctx.size(), g->d_merged, nullptr, c0->stream));
  c0->launches++;
  g->merged_valid = true;
  return RQCD_OK;
}

int read_merged(rqcd_group* g, rqcd_summary* out) {
  rqcd_ctx* c0 = g->ctx[0];
  DeviceGuard guard(c0->device);
  RQCD_GCU(cudaMemcpyAsync(g->h_merged, g->d_merged, sizeof(DevSummary), cudaMemcpyDeviceToHost, c0->stream));
  RQCD_GCU(cudaStreamSynchronize(c0->stream));
  memcpy(out, g->h_merged, sizeof *out);
  return depth_cap_status(*out);
}

}  // namespace

extern "C" {

void rqcd_shard_range(int64_t n, int32_t k, int32_t n_shards, int64_t* lo, int64_t* hi) {
  if (n_shards < 1) n_shards = 1;
  // k*n/n_shards without overflowing 64 bits for the sizes at hand (n < 2^55, n_shards <= 256)
  if (lo) *lo = (int64_t)((__int128)n * k / n_shards);
  if (hi) *hi = (int64_t)((__int128)n * (k + 1) / n_shards);
}

int rqcd_group_create(rqcd_group** out, const int32_t* devices, int32_t n_devices) {
  if (!out) return RQCD_ERR_INVALID_ARG;
  *out = nullptr;
  if (!devices || n_devices < 1 || n_devices > 256) return RQCD_ERR_INVALID_ARG;
  rqcd_group* g = new (std::nothrow) rqcd_group();
  if (!g) return RQCD_ERR_NOMEM;
  for (int k = 0; k < n_devices; k++) {
    rqcd_ctx* c = nullptr;
    int st = rqcd_create(&c, devices[k]);
    if (st != RQCD_OK) {
      rqcd_group_destroy(g);
      return st;
    }
    g->ctx.push_back(c);
  }
  const int d0 = g->ctx[0]->device;
  // can every other device store into the first one's memory?
  for (int k = 1; k < n_devices && g->peer_slots; k++) {
    const int dk = g->ctx[k]->device;
    if (dk == d0) continue;
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, dk, d0) != cudaSuccess || !can) g->peer_slots = false;
  }
  cudaGetLastError();
  const char* force = getenv("RQCD_GROUP_HOST_SLOTS");
  if (force && force[0] == '1') g->peer_slots = false;
  cudaError_t e = cudaSuccess;
  {
    DeviceGuard guard(d0);
    if (g->peer_slots) e = cudaMalloc(&g->slots, sizeof(DevSummary) * n_devices);
    else e = cudaHostAlloc(&g->slots, sizeof(DevSummary) * n_devices, cudaHostAllocPortable | cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaMalloc(&g->d_merged, sizeof(DevSummary));
    if (e == cudaSuccess) e = cudaMallocHost(&g->h_merged, sizeof(DevSummary));
  }
  for (int k = 0; k < n_devices && e == cudaSuccess; k++) {
    const int dk = g->ctx[k]->device;
    DeviceGuard guard(dk);
    if (g->peer_slots && dk != d0) {
      e = cudaDeviceEnablePeerAccess(d0, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        e = cudaSuccess;
      }
    }
    cudaEvent_t ev = nullptr;
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    g->published.push_back(ev);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    rqcd_group_destroy(g);
    return RQCD_ERR_CUDA;
  }
  *out = g;
  return RQCD_OK;
}

void rqcd_group_destroy(rqcd_group* g) {
  if (!g) return;
  for (rqcd_ctx* c : g->ctx) rqcd_synchronize(c);
  for (size_t k = 0; k < g->published.size(); k++)
    if (g->published[k]) {
      DeviceGuard guard(g->ctx[k]->device);
      cudaEventDestroy(g->published[k]);
    }
  if (!g->ctx.empty()) {
    DeviceGuard guard(g->ctx[0]->device);
    if (g->slots) {
      if (g->peer_slots) cudaFree(g->slots);
      else cudaFreeHost(g->slots);
    }
    if (g->d_merged) cudaFree(g->d_merged);
    if (g->h_merged) cudaFreeHost(g->h_merged);
  }
  for (rqcd_ctx* c : g->ctx) rqcd_destroy(c);
  delete g;
}

int32_t rqcd_group_size(const rqcd_group* g) { return g ? (int32_t)g->ctx.size() : 0; }

rqcd_ctx* rqcd_group_ctx(rqcd_group* g, int32_t k) {
  return (g && k >= 0 && k < (int32_t)g->ctx.size()) ? g->ctx[k] : nullptr;
}

const char* rqcd_group_last_error(const rqcd_group* g) { return g ? g->last_error : ""; }

int rqcd_group_peer_slots(const rqcd_group* g) { return g && g->peer_slots ? 1 : 0; }

int rqcd_group_set_obstacles(rqcd_group* g, const rqcd_obstacle* obstacles, int32_t m) {
  if (!g) return RQCD_ERR_INVALID_ARG;
  for (rqcd_ctx* c : g->ctx) {
    int st = rqcd_set_obstacles(c, obstacles, m);
    if (st != RQCD_OK) return group_fail(g, st, "rqcd_set_obstacles", c);
  }
  return RQCD_OK;
}

int rqcd_group_generate_check(rqcd_group* g, const rqcd_bc_soa* in, int64_t n, double min_time_section, uint32_t flags,
                              rqcd_out* out) {
  if (!g || !in || !out || n < 0 || (flags & RQCD_F_DEVICE_PTRS)) return RQCD_ERR_INVALID_ARG;
  if ((n > 0 && !in->data) || in->first < 0 || in->stride < in->first + n) return RQCD_ERR_INVALID_ARG;
  const int G = (int)g->ctx.size();
  const int m = g->ctx[0]->m;
  std::vector<int> status(G, RQCD_OK);
  auto shard = [&](int k) {
    int64_t lo, hi;
    rqcd_shard_range(n, k, G, &lo, &hi);
    rqcd_bc_soa sin = *in;
    sin.first = in->first + lo;  // the same matrix: per-candidate rows from element first + lo, shared rows by group
    rqcd_out so = *out;
    so.summary = nullptr;
    if (so.verdict) so.verdict += lo;
    if (so.first_obstacle) so.first_obstacle += lo;
    if (so.verdict_matrix) so.verdict_matrix += lo * (int64_t)m;
    if (so.cost) so.cost += lo;
    if (so.coeffs) so.coeffs += lo;
    status[k] = rqcd_generate_check(g->ctx[k], &sin, hi - lo, min_time_section, flags, lo, &so);
    if (status[k] == RQCD_OK || status[k] == RQCD_ERR_DEPTH_CAP) {
      int st = publish(g, k);
      if (st != RQCD_OK) status[k] = st;
    }
  };
  // one host thread per shard: each drives its own device's streaming pipeline (chunked H2D beside the persistent kernel)
  std::vector<std::thread> workers;
  for (int k = 1; k < G; k++) workers.emplace_back(shard, k);
  shard(0);
  for (std::thread& t : workers) t.join();
  for (int k = 0; k < G; k++)
    if (status[k] != RQCD_OK && status[k] != RQCD_ERR_DEPTH_CAP) return group_fail(g, status[k], "rqcd_generate_check (shard)", g->ctx[k]);
  int st = merge(g);
  if (st != RQCD_OK) return st;
  rqcd_summary merged;
  st = read_merged(g, &merged);
  if (out->summary) *out->summary = merged;
  return st;
}

int rqcd_group_generate_check_resident(rqcd_group* g, const rqcd_bc_soa* in, const int64_t* n, double min_time_section,
                                       uint32_t flags, int64_t index_base, const rqcd_out* out, rqcd_summary* merged) {
  if (!g || !in || !n) return RQCD_ERR_INVALID_ARG;
  const int G = (int)g->ctx.size();
  int64_t base = index_base;
  for (int k = 0; k < G; k++) {
    rqcd_out so = {};
    if (out) so = out[k];
    so.summary = nullptr;
    int st = rqcd_generate_check(g->ctx[k], &in[k], n[k], min_time_section, flags | RQCD_F_DEVICE_PTRS, base, &so);
    if (st != RQCD_OK) return group_fail(g, st, "rqcd_generate_check (shard)", g->ctx[k]);
    st = publish(g, k);
    if (st != RQCD_OK) return st;
    base += n[k];
  }
  int st = merge(g);
  if (st != RQCD_OK) return st;
  return merged ? read_merged(g, merged) : RQCD_OK;
}

int rqcd_group_read_summary(rqcd_group* g, rqcd_summary* merged) {
  if (!g || !merged || !g->merged_valid) return RQCD_ERR_INVALID_ARG;
  return read_merged(g, merged);
}

// ---- one rank per process ------------------------------------------------------------------------------------------
int rqcd_comm_unique_id(void* id) {
  if (!id) return RQCD_ERR_INVALID_ARG;
  NcclApi* a = nccl_api();
  if (!a->handle) return RQCD_ERR_UNSUPPORTED;
  ncclUniqueId u;
  if (a->GetUniqueId(&u) != ncclSuccess) return RQCD_ERR_CUDA;
  memcpy(id, &u, sizeof u);
  return RQCD_OK;
}

int rqcd_comm_init(rqcd_ctx* ctx, const void* id, int32_t rank, int32_t world) {
  if (!ctx || !id || world < 1 || rank < 0 || rank >= world || ctx->comm) return RQCD_ERR_INVALID_ARG;
  NcclApi* a = nccl_api();
  if (!a->handle) {
    snprintf(ctx->last_error, sizeof ctx->last_error, "%s", a->error);
    return RQCD_ERR_UNSUPPORTED;
  }
  DeviceGuard guard(ctx->device);
  ncclUniqueId u;
  memcpy(&u, id, sizeof u);
  ncclComm_t comm = nullptr;
  ncclResult_t r = a->CommInitRank(&comm, world, u, rank);
  if (r != ncclSuccess) return fail_nccl(ctx, a, r, "ncclCommInitRank");
  if (ctx->d_gather) cudaFree(ctx->d_gather);
  ctx->d_gather = nullptr;
  cudaError_t e = cudaMalloc(&ctx->d_gather, sizeof(DevSummary) * world);
  if (e != cudaSuccess) {
    a->CommDestroy(comm);
    return fail_cuda(ctx, e, "cudaMalloc(gather table)");
  }
  ctx->comm = comm;
  ctx->comm_rank = rank;
  ctx->comm_world = world;
  return RQCD_OK;
}

int rqcd_comm_finalize(rqcd_ctx* ctx) {
  if (!ctx) return RQCD_ERR_INVALID_ARG;
  if (!ctx->comm) return RQCD_OK;
  DeviceGuard guard(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  nccl_api()->CommDestroy((ncclComm_t)ctx->comm);
  ctx->comm = nullptr;
  ctx->comm_rank = 0;
  ctx->comm_world = 1;
  if (ctx->d_gather) cudaFree(ctx->d_gather);
  ctx->d_gather = nullptr;
  return RQCD_OK;
}

int32_t rqcd_comm_world(const rqcd_ctx* ctx) { return ctx ? ctx->comm_world : 0; }

int rqcd_exchange_summaries(rqcd_ctx* ctx, rqcd_summary* d_record) {
  if (!ctx || !ctx->summary_valid) return RQCD_ERR_INVALID_ARG;
  DeviceGuard guard(ctx->device);
  if (!ctx->d_merged) RQCD_CU(cudaMalloc(&ctx->d_merged, sizeof(DevSummary)));
  const DevSummary* table = ctx->d_summary;
  if (ctx->comm) {
    NcclApi* a = nccl_api();
    ncclResult_t r = a->AllGather(ctx->d_summary, ctx->d_gather, sizeof(DevSummary), ncclChar, (ncclComm_t)ctx->comm, ctx->stream);
    if (r != ncclSuccess) return fail_nccl(ctx, a, r, "ncclAllGather");
    table = ctx->d_gather;
  }
  RQCD_CU(launch_merge_summaries(table, ctx->comm_world, ctx->d_merged, reinterpret_cast<DevSummary*>(d_record), ctx->stream));
  ctx->launches++;
  ctx->merged_valid = true;
  return RQCD_OK;
}

int rqcd_read_merged_summary(rqcd_ctx* ctx, rqcd_summary* merged) {
  if (!ctx || !merged || !ctx->merged_valid) return RQCD_ERR_INVALID_ARG;
  DeviceGuard guard(ctx->device);
  RQCD_CU(cudaMemcpyAsync(ctx->h_summary, ctx->d_merged, sizeof(DevSummary), cudaMemcpyDeviceToHost, ctx->stream));
  RQCD_CU(cudaStreamSynchronize(ctx->stream));
  memcpy(merged, ctx->h_summary, sizeof *merged);
  return depth_cap_status(*merged);
}

}  // extern "C"
