// rqcd_kernel_util.cuh -- helpers shared by the kernel translation units (rqcd_kernels.cu, rqcd_pool.cu): TMA bulk copy
// of the obstacle table, the streaming-input wait, trajectory construction for one lane, partial reduction.
#pragma once
#include "rqcd_device.cuh"
#include "rqcd_kernels.cuh"

namespace rqcd {
namespace {

constexpr unsigned kFull = 0xffffffffu;


// ---- TMA bulk copy global -> shared (1-D), completion on an mbarrier ---------------------------------
RQCD_DEV uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

RQCD_DEV void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
RQCD_DEV void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
RQCD_DEV void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "RQCD_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra RQCD_DONE;\n"
      "bra RQCD_WAIT;\n"
      "RQCD_DONE:\n"
      "}\n" ::"r"(smem_addr(bar)),
      "r"(parity)
      : "memory");
}


RQCD_DEV bool better(double ca, long long ia, double cb, long long ib) {
  // candidate a beats b: lower cost, or equal cost and lower index; an empty slot has index -1
  if (ia < 0) return false;
  if (ib < 0) return true;
  return ca < cb || (ca == cb && ia < ib);
}

// Inputs are read once and may still be arriving over PCIe while the kernel runs (CheckParams::progress): load
// them at L2 (ld.global.cg), the coherence point with the copy engine, never through a possibly stale L1 line.
RQCD_DEV double ld_in(const double* p) { return __ldcg(p); }

// Streaming input (host-pointer calls): block until the copy engine has delivered candidates [0, want). All 32 lanes
// call it together. Lane 0 polls the progress word with an ACQUIRE load at system scope -- the copy engine wrote the
// rows before it advanced the word -- and the warp barrier orders the other lanes' loads behind it; the answer is
// warp-uniform. A copy that never comes (bounded by CheckParams::stream_timeout cycles) fails the call instead of
// hanging the GPU: *stream_error is set and false returned.
RQCD_DEV bool wait_for_input(const CheckParams& p, unsigned long long want, unsigned long long& avail, int lane) {
  if (want <= avail) return true;  // warp-uniform: `avail` is
  int ok = 1;
  if (lane == 0) {
    const long long t_begin = clock64();
    for (;;) {
      unsigned long long v;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p.progress) : "memory");
      avail = v;
      if (avail >= want) break;
      if (clock64() - t_begin > p.stream_timeout) {
        atomicExch(p.stream_error, 1ull);
        ok = 0;
        break;
      }
      __nanosleep(256);
    }
  }
  avail = __shfl_sync(kFull, avail, 0);
  ok = __shfl_sync(kFull, ok, 0);
  __syncwarp();
  return ok != 0;
}

// ---- boundary-condition rows (rqcd_bc_soa) ---------------------------------------------------------------
// Where candidate i of the launch finds its element: in a per-candidate row, in a shared (per-group) row.
struct BcIndex {
  long long cand, grp;
};
RQCD_DEV BcIndex bc_index(const CheckParams& p, long long i) {
  BcIndex x;
  x.cand = i + p.in_first;
  x.grp = 0;
  if (p.shared_rows != 0u) {
    const unsigned long long q = (unsigned long long)(i + p.grp_first), g = (unsigned long long)p.group_size;
    x.grp = (long long)(((q | g) >> 32) == 0 ? (unsigned long long)((unsigned)q / (unsigned)g) : q / g);
  }
  return x;
}
RQCD_DEV double bc_at(const CheckParams& p, int r, const BcIndex& x) {
  return ld_in(p.in + r * p.in_stride + (((p.shared_rows >> r) & 1u) ? x.grp : x.cand));
}

// ---- trajectory construction for one lane ------------------------------------------------------------
template <int MODE>
RQCD_DEV void load_trajectory(const CheckParams& p, long long i, double (&c)[18], double& t0, double& t1,
                              double& cost) {
  if (MODE == MODE_BC) {
    const BcIndex x = bc_index(p, i);
    const double Tf = bc_at(p, 18, x);
    double total = 0.0;
#pragma unroll
    for (int ax = 0; ax < 3; ax++) {
      const double p0 = bc_at(p, 0 + ax, x), v0 = bc_at(p, 3 + ax, x), a0 = bc_at(p, 6 + ax, x);
      const double pf = bc_at(p, 9 + ax, x), vf = bc_at(p, 12 + ax, x), af = bc_at(p, 15 + ax, x);
      const AxisParams q = gen_axis(p0, v0, a0, pf, vf, af, (p.goal_mask >> (3 * ax)) & 1,
                                    (p.goal_mask >> (3 * ax + 1)) & 1, (p.goal_mask >> (3 * ax + 2)) & 1, Tf);
      axis_coeffs(q, p0, v0, a0, c[0 + ax], c[3 + ax], c[6 + ax], c[9 + ax], c[12 + ax], c[15 + ax]);
      total = (ax == 0) ? q.cost : total + q.cost;  // GetCost: cost[0] + cost[1] + cost[2] (.hpp:214-216)
    }
    cost = total;
    t0 = 0.0;
    t1 = Tf;
  } else {
#pragma unroll
    for (int j = 0; j < 18; j++) c[j] = ld_in(p.in + j * p.in_stride + i);
    t0 = p.t_start ? ld_in(p.t_start + i) : 0.0;
    t1 = ld_in(p.t_end + i);
    cost = p.cost_in ? ld_in(p.cost_in + i) : __longlong_as_double(0x7ff8000000000000ll);
  }
}


}  // namespace
}  // namespace rqcd
