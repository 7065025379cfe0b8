// rqcd_kernels.cu -- CUDA kernels of the hot path for sm_100a. Compile with -fmad=false (see rqcd_device.cuh).
//
// k_check is the fused kernel: RapidTrajectoryGenerator::Generate -> GetTrajectory ->
// CollisionChecker::CollisionCheck against the whole obstacle set, i.e. the candidate-loop body of
// test/ForestPerformanceEval.cpp:186-225 (and, in pairwise form, test/PerformanceEval.cpp:140-182).
//
// Shape of the kernel
//   * persistent CTAs (grid = SMs x resident CTAs); the prepared obstacle table is staged once per CTA into
//     shared memory with one TMA bulk copy (cp.async.bulk + mbarrier) and reused by every candidate.
//   * a lane owns one trajectory (its 18 coefficients live in registers) and walks the obstacles on its own:
//     every trip of the warp loop runs exactly one CollisionCheckSection frame per lane, so a lane that
//     needs 15 frames for one obstacle does not hold back its 31 neighbours -- they move on to their next
//     obstacles. The reference's recursion is a per-lane explicit stack of pending [ts, tf] intervals.
//   * the frame is one converged instruction stream for the whole warp (see frame() in rqcd_device.cuh): the
//     only divergent region is the transcendental core of the cubic solver. The only position a frame evaluates is
//     X(mid): the plane-side tests at the critical times and at the interval's ends are the sign of a scalar
//     quintic (5 FMAs each) with a certified error margin and a literal fallback, so an interval is just [ts, tf].
//     Divisions and square roots are nvcc's in-line sequences with shared reciprocal refinements and ONE range
//     vote per frame (FastMath / PlainMath).
//   * CollisionCheck's end-point tests (IsPointInside of X(t_start), X(t_end)) are answered for the whole table when
//     a trajectory is loaded -- lane = obstacle, one ballot per 32 obstacles -- for tables of 16+ obstacles; shorter
//     tables, moving obstacles and the pairwise shape keep them in the pair's first frame.
//   * work queue: when enough lanes have finished their trajectory the warp takes new candidates from a global
//     counter (ballot + popc, one atomicAdd per warp). A few new candidates are built by the whole warp (one lane
//     per axis), many by their own lanes. With host buffers the candidates may still be arriving over PCIe: a warp
//     then waits for the copy engine's progress word.
//   * CTA lockstep (CheckParams::lockstep, chosen per launch): one __syncthreads_count per trip keeps the warps of a
//     CTA in the same part of the loop body and aligns their fetches -- used for short tables and the pairwise
//     shape; tables of 16+ obstacles run with free-running warps.
//   * per-CTA partial verdict counts / best candidate, reduced in block order by k_finalize (deterministic).
//
// k_check is the kernel of the PAIRWISE shape and of SHORT static tables (and of every shape with RQCD_POOL=0); obstacle
// tables go to k_pool (rqcd_pool.cu: staged sections, certified solver skip, tasks decoupled from their owner lanes).
// Same device functions, same verdicts. Staged sections were tried here too (PLANE / SOLVE trips per warp) and cost 20 %:
// a lane that owns its trajectory idles while it waits for its kind of trip.
// This file also holds k_feas (CheckInputFeasibility, sections decided on squares), the accept list of the input filter,
// k_planes, k_generate, the solver unit kernels, k_fragile, k_finalize and the multi-GPU summary kernels.
#include <cstdio>
#include "rqcd_kernel_util.cuh"

namespace rqcd {

namespace {

constexpr int kCoopOwners = 10;  // new candidates a warp builds cooperatively in one pass (3 lanes each)

// Per-trajectory constants (window, cost) are read once per obstacle: they live in a volatile per-thread local
// array (L1-resident), not in registers.
enum TrajRow { SLOT_T0 = 0, SLOT_T1, SLOT_COST, kTrajRows };
constexpr int kSlotRows = EndPoints::kRows;  // per-thread shared-memory rows

#ifdef RQCD_MAXNREG  // kernel-variant experiments: an explicit register cap instead of a resident-CTA target
#define RQCD_CHECK_BOUNDS __maxnreg__(RQCD_MAXNREG)
#else
#define RQCD_CHECK_BOUNDS __launch_bounds__(kThreads, RQCD_MIN_BLOCKS)
#endif
template <int MODE, bool MOVING, bool PAIRWISE, bool ENDBITS>
__global__ void RQCD_CHECK_BOUNDS k_check(const CheckParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m = PAIRWISE ? 1 : p.m, mp = p.mp;

  double* tab = reinterpret_cast<double*>(smem_raw);
  int* oflags = reinterpret_cast<int*>(tab + (size_t)p.stride * mp);
  unsigned long long* tcnt = reinterpret_cast<unsigned long long*>(oflags + mp);  // [8][kThreads]: verdict counters
  double* bcost = reinterpret_cast<double*>(tcnt + 8 * kThreads);                 // [kThreads]
  long long* bidx = reinterpret_cast<long long*>(bcost + kThreads);               // [kThreads]
  double* slot = reinterpret_cast<double*>(bidx + kThreads);                      // [kSlotRows][kThreads]: EndPoints rows
  double* base = slot + kSlotRows * kThreads;                                     // MOVING: [18][kThreads]
  // static table: bit k of a thread's words = an end point of its trajectory is inside obstacle k (kEndBits)
  constexpr bool kEndBits = ENDBITS && !MOVING && !PAIRWISE;
  unsigned* inbits = reinterpret_cast<unsigned*>(base);                            // [(mp + 31) / 32][kThreads]
  __shared__ __align__(8) uint64_t mbar;
  volatile double traj_const[kTrajRows];
#define RQCD_SLOT(row) traj_const[row]

  const uint32_t blob_bytes = (uint32_t)(((size_t)p.stride * mp) * 8 + (size_t)mp * 4);
  if (tid == 0) mbar_init(&mbar, 1);
#pragma unroll
  for (int j = 0; j < 8; j++) tcnt[j * kThreads + tid] = 0ull;
  bcost[tid] = __longlong_as_double(0x7ff0000000000000ll);
  bidx[tid] = -1;
  __syncthreads();
  if (blob_bytes > 0 && !PAIRWISE) {
    if (tid == 0) tma_bulk_g2s(tab, p.obs_blob, blob_bytes, &mbar);
    mbar_wait(&mbar, 0);
  }

  // ---- lane state: the trajectory (coefficients in registers for its whole life), the current pair's
  // interval [ts, tf] with the positions at its ends, and the stack of parked low children
  bool have = false, in_pair = false, exhausted = false, first_frame = false;
  long long idx = 0;
  double c[18];
#pragma unroll
  for (int j = 0; j < 18; j++) c[j] = 0.0;
  double ts = 0, tf = 0;
  EndPoints ends;
  ends.slot = slot + tid;
  ends.stride = kThreads;
  ends.set_xs(v3(0, 0, 0));
  ends.set_xf(v3(0, 0, 0));
  int k = 0, verdict = 0, first = -1, sp = 0;
  unsigned pair_hits = 0, depth_cap_hits = 0;  // this trajectory's pairs: Collision | Indeterminable << 16; verdict 3
  double2 stack[kMaxDepth];  // parked low children {ts, lo_tf}
  LaneSphere sph;
  sph.cx = sph.cy = sph.cz = sph.r = 0;
  const unsigned lt_mask = (1u << lane) - 1u;

  int n_have = 0;
  unsigned long long avail = 0;  // candidates known to have arrived (streaming input)
  // with an accept list (input filter) the work items are positions of the list, else the candidates themselves
  const long long n_work = p.n_list ? (long long)*p.n_list : p.n;
  for (;;) {
    // ---- CTA lockstep (p.lockstep, chosen per launch): one barrier per trip keeps the warps of a CTA in the same
    // part of the loop body, which is larger than the instruction cache, so an instruction line fetched from L2
    // serves all of them; the same instruction counts the CTA's idle lanes for the work-queue trigger. Long obstacle
    // tables (few fetches, long uninterrupted runs of frames) do better with free-running warps.
    bool fetch_now = true;
    if (p.lockstep) {
      n_have = __syncthreads_count(have);
      fetch_now = (kThreads - n_have >= p.refill_min * (kThreads / 32)) || n_have == 0;
    }
    // ---- work queue: idle lanes take new candidates
    const unsigned need = __ballot_sync(kFull, !have);
    if (need != 0 && !exhausted && fetch_now) {
      const int cnt = __popc(need);
      if (p.lockstep || cnt >= p.refill_min || need == kFull) {
        unsigned long long b = 0;
        if (lane == 0) b = atomicAdd(p.work_counter, (unsigned long long)cnt);
        b = __shfl_sync(kFull, b, 0);
        if (b + (unsigned long long)cnt >= (unsigned long long)n_work) exhausted = true;
        bool arrived = true;
        if (p.progress) {  // streaming input: wait until the copy engine has delivered everything this warp took
          unsigned long long want = b + (unsigned long long)cnt;
          if (want > (unsigned long long)p.n) want = (unsigned long long)p.n;
          arrived = wait_for_input(p, want, avail, lane);
          if (!arrived) exhausted = true;
        }
        bool fresh = false;
        const long long my_pos = (long long)b + __popc(need & lt_mask);
        const long long my_i = (p.list && my_pos < n_work) ? p.list[my_pos] : my_pos;
        // planner: skip[i] != 0 was dropped by an earlier stage
        const bool take = !have && my_pos < n_work && arrived && !(p.skip && p.skip[my_i] != 0);
        double t0 = 0.0, t1 = 0.0, cost = 0.0;
        V3 x0 = v3(0, 0, 0), x1 = v3(0, 0, 0);
        double amax = 0.0;
        const unsigned owners = __ballot_sync(kFull, take);
        const int n_own = __popc(owners);
        if (MODE == MODE_BC && n_own <= kCoopOwners) {
          // RapidTrajectoryGenerator::Generate + GetTrajectory for a FEW new candidates by the whole warp: the three
          // axes of a primitive are independent (SingleAxisTrajectory), so lane 3q + a builds axis a of the q-th new
          // candidate and the owner collects its 18 coefficients with shuffles. Same arithmetic per axis as
          // load_trajectory; a handful of idle lanes no longer walk three axes one after the other while 28 wait.
          const int my_rank = __popc(owners & lt_mask);
          const int q = lane / 3, ax = lane % 3;
          double tc[6] = {0, 0, 0, 0, 0, 0}, t_cost = 0.0, t_x0 = 0.0, t_x1 = 0.0, t_am = 0.0, t_tf = 0.0;
          const bool helper = lane < 3 * kCoopOwners && q < n_own;
          const unsigned own = helper ? __fns(owners, 0, q + 1) : 0u;
          const long long own_i = __shfl_sync(kFull, my_i, (int)own);  // the q-th owner's candidate
          if (helper) {
            const long long i = own_i;
            const BcIndex x = bc_index(p, i);
            const double Tf = bc_at(p, 18, x);
            const double p0 = bc_at(p, 0 + ax, x), v0 = bc_at(p, 3 + ax, x), a0 = bc_at(p, 6 + ax, x);
            const double pf = bc_at(p, 9 + ax, x), vf = bc_at(p, 12 + ax, x), af = bc_at(p, 15 + ax, x);
            const AxisParams g = gen_axis(p0, v0, a0, pf, vf, af, (p.goal_mask >> (3 * ax)) & 1,
                                          (p.goal_mask >> (3 * ax + 1)) & 1, (p.goal_mask >> (3 * ax + 2)) & 1, Tf);
            axis_coeffs(g, p0, v0, a0, tc[0], tc[1], tc[2], tc[3], tc[4], tc[5]);
            t_cost = g.cost;
            t_tf = Tf;
            if (p.coeffs) {
#pragma unroll
              for (int j = 0; j < 6; j++) p.coeffs[(3 * j + ax) * p.coeff_stride + i] = tc[j];
            }
            // this axis' share of X(t_start = 0), X(t_end = Tf) and of the side filter's bound
            t_x0 = poly5(tc[0], tc[1], tc[2], tc[3], tc[4], tc[5], 0.0);
            t_x1 = poly5(tc[0], tc[1], tc[2], tc[3], tc[4], tc[5], Tf);
            const double T = fabs(Tf);
#pragma unroll
            for (int j = 0; j < 6; j++) t_am = t_am * T + fabs(tc[j]);
          }
          __syncwarp();
          const int src = take ? 3 * my_rank : 0;
#pragma unroll
          for (int j = 0; j < 6; j++) {
            const double ca = __shfl_sync(kFull, tc[j], src), cb = __shfl_sync(kFull, tc[j], src + 1);
            const double cc = __shfl_sync(kFull, tc[j], src + 2);
            if (take) c[3 * j] = ca, c[3 * j + 1] = cb, c[3 * j + 2] = cc;
          }
          const double k0 = __shfl_sync(kFull, t_cost, src), k1 = __shfl_sync(kFull, t_cost, src + 1);
          const double k2 = __shfl_sync(kFull, t_cost, src + 2);
          x0 = v3(__shfl_sync(kFull, t_x0, src), __shfl_sync(kFull, t_x0, src + 1), __shfl_sync(kFull, t_x0, src + 2));
          x1 = v3(__shfl_sync(kFull, t_x1, src), __shfl_sync(kFull, t_x1, src + 1), __shfl_sync(kFull, t_x1, src + 2));
          amax = __shfl_sync(kFull, t_am, src) + __shfl_sync(kFull, t_am, src + 1) + __shfl_sync(kFull, t_am, src + 2);
          t1 = __shfl_sync(kFull, t_tf, src);
          cost = k0 + k1 + k2;  // GetCost: cost[0] + cost[1] + cost[2] (.hpp:214-216)
          if (take && p.cost) p.cost[my_i] = cost;
        } else if (take) {
          load_trajectory<MODE>(p, my_i, c, t0, t1, cost);
          if (MODE == MODE_BC) {
            if (p.cost) p.cost[my_i] = cost;
            if (p.coeffs) {
#pragma unroll
              for (int j = 0; j < 18; j++) p.coeffs[j * p.coeff_stride + my_i] = c[j];
            }
          }
          if (!MOVING) {
            x0 = traj_value(c, t0);
            x1 = traj_value(c, t1);
            amax = side_filter_bound(c, t0, t1);
          }
        }
        if (take) {
          {
            const long long i = my_i;
            idx = i;
            RQCD_SLOT(SLOT_T0) = t0;
            RQCD_SLOT(SLOT_T1) = t1;
            RQCD_SLOT(SLOT_COST) = cost;
            if (MOVING) {
#pragma unroll
              for (int j = 0; j < 18; j++) base[j * kThreads + tid] = c[j];
            } else {
              // X(t_start), X(t_end) of the trajectory and the side filter's bound: the same for every obstacle
              ends.set_xs(x0);
              ends.set_xf(x1);
              ends.set_amax(amax);
            }
            if (PAIRWISE) {
              sph.cx = ld_in(p.spheres + i);
              sph.cy = ld_in(p.spheres + p.sphere_stride + i);
              sph.cz = ld_in(p.spheres + 2 * p.sphere_stride + i);
              sph.r = ld_in(p.spheres + 3 * p.sphere_stride + i);
            }
            have = true;
            fresh = true;
            in_pair = false;
            k = 0;
            verdict = 0;
            first = -1;
          }
        }
        if (kEndBits) {
          // CollisionCheck's end-point tests (CollisionChecker.cpp:73-77) for the new trajectories against the whole
          // table, by the whole warp: lane j tests obstacles j, j + 32, ... against the end points of one new
          // trajectory at a time (read from its owner's rows), a ballot packs 32 answers into a word for the owner.
          unsigned todo = __ballot_sync(kFull, fresh);
          __syncwarp();
          while (todo) {
            const int owner = __ffs(todo) - 1;
            todo &= todo - 1;
            const double* row = slot + (tid - lane + owner);
            const V3 Xs = v3(row[0], row[kThreads], row[2 * kThreads]);
            const V3 Xf = v3(row[3 * kThreads], row[4 * kThreads], row[5 * kThreads]);
            for (int k0 = 0; k0 < m; k0 += 32) {
              const int kk = k0 + lane;
              bool inside = false;
              if (kk < m) {
                TableObs o;
                o.rec = tab + (size_t)kk * p.stride;
                inside = ends_inside(o, Xs, Xf);
              }
              const unsigned bits = __ballot_sync(kFull, inside);
              if (lane == owner) inbits[(k0 >> 5) * kThreads + tid] = bits;
            }
          }
        }
      }
    }
    if (p.lockstep) {
      // nothing anywhere in the CTA, the fetch above brought nothing and every warp has seen the queue run dry (a
      // fetch can also come back empty-handed when the planner's earlier stages rejected all it took)
      if (n_have == 0 && __syncthreads_count(have || !exhausted) == 0) break;
      if (__ballot_sync(kFull, have) == 0) continue;  // this warp is done; keep meeting the CTA at the barrier
    } else {
      if (exhausted && __ballot_sync(kFull, have) == 0) break;
    }

    int pv = -1;  // verdict of the pair that completed in this trip, if any

    // ---- CollisionChecker::CollisionCheck entry for the lane's next obstacle (CollisionChecker.cpp:70-80);
    // its end-point tests ride on the first frame below
    if (kEndBits && have && !in_pair && k < m && ((inbits[(k >> 5) * kThreads + tid] >> (k & 31)) & 1u)) {
      // pairs whose end-point test is positive are Collisions without a section (CollisionChecker.cpp:73-77)
      while (k < m && ((inbits[(k >> 5) * kThreads + tid] >> (k & 31)) & 1u) && !(p.early_exit && verdict != 0)) {
        if (p.verdict_matrix) p.verdict_matrix[idx * (long long)m + k] = (uint8_t)1;
        if (first < 0) {
          verdict = 1;
          first = k;
        }
        pair_hits += 1u;
        k++;
      }
    }
    if (have && !in_pair && k < m && !(kEndBits && p.early_exit && verdict != 0)) {
      ts = RQCD_SLOT(SLOT_T0);
      tf = RQCD_SLOT(SLOT_T1);
      in_pair = true;
      first_frame = true;
      sp = 0;
      if (MOVING) {
        // Trajectory::operator- (Trajectory.hpp:121-128): coefficient-wise difference on the common window
        const int fl = oflags[k];
        if (fl & OBS_MOVING) {
          const double* rec = tab + (size_t)k * p.stride;
#pragma unroll
          for (int j = 0; j < 18; j++) c[j] = base[j * kThreads + tid] - rec[OF_MOTION + j];
          const double ms = rec[OF_MT0], me = rec[OF_MT1];
          const double t0 = ts, t1 = tf;
          ts = (t0 < ms) ? ms : t0;  // std::max(_startTime, rhs start)
          tf = (me < t1) ? me : t1;  // std::min(_endTime, rhs end)
          if (!(ts <= tf)) {  // no common time (the reference would assert, Trajectory.hpp:68): nothing to hit
            pv = 0;
            in_pair = false;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 18; j++) c[j] = base[j * kThreads + tid];
        }
        ends.set_xs(traj_value(c, ts));
        ends.set_xf(traj_value(c, tf));
        ends.set_amax(side_filter_bound(c, ts, tf));
      }
    }

    // ---- one CollisionCheckSection frame (CollisionChecker.cpp:82-193), executed by the whole warp: lanes
    // without a pair run it on stale data and drop the result, which keeps the warp converged
    {
      Children ch;
      int r;
      if (PAIRWISE) {
        r = frame<true>(c, ts, tf, ends, first_frame, in_pair, sph, p.min_t, ch);
      } else {
        TableObs o;
        const int kk = min(k, max(m - 1, 0));
        o.rec = tab + (size_t)kk * p.stride;
        r = frame<!kEndBits>(c, ts, tf, ends, first_frame, in_pair, o, p.min_t, ch);
      }
      first_frame = false;
      // state update, written with selects: one instruction stream for the common cases (no child; one child)
      const bool child = in_pair & (r == STEP_CHILD);
      const bool to_hi = child & ch.has_hi, to_lo = child & !ch.has_hi;
      const bool clear = in_pair & (r == STEP_CLEAR);
      if (to_hi & ch.has_lo) {  // both children: the parent's low tail call waits until the high subtree is clear
        if (sp < kMaxDepth) {
          stack[sp++] = make_double2(ts, ch.lo_tf);
        } else {
          pv = 3;
          in_pair = false;
        }
      }
      ts = to_hi ? ch.t_next : ts;
      tf = to_lo ? ch.t_next : tf;
      if (clear & (sp > 0)) {  // subtree clear: resume the innermost parked low child
        const double2 e = stack[--sp];
        ts = e.x;
        tf = e.y;
      } else if (in_pair & !child) {  // NoCollision with nothing parked, Collision or Indeterminable: the pair is done
        pv = clear ? 0 : r;
        in_pair = false;
      }
    }

    // ---- bookkeeping. A finished pair only bumps the lane's packed per-trajectory counters; the CTA counters in
    // shared memory are touched once per finished trajectory
    if (pv >= 0) {
      if (p.verdict_matrix) p.verdict_matrix[idx * (long long)m + k] = (uint8_t)pv;
      if (pv != 0 && first < 0) {
        verdict = pv;
        first = k;
      }
      pair_hits += (pv == 1 ? 1u : 0u) + (pv == 2 ? 0x10000u : 0u);
      depth_cap_hits += (pv == 3);
      k++;
    }
    const bool fin = have && !in_pair && (k >= m || (p.early_exit && verdict != 0));
    {
      if (fin) {
        // the finishing lane adds its trajectory to its own counter rows (summed over the CTA at the end)
        const unsigned n1 = pair_hits & 0xffffu, n2 = pair_hits >> 16;  // k <= m < 2^16
        unsigned long long* t = tcnt + tid;
        t[verdict * kThreads] += 1ull;
        t[4 * kThreads] += (unsigned)k - n1 - n2 - depth_cap_hits;
        t[5 * kThreads] += n1;
        t[6 * kThreads] += n2;
        if (depth_cap_hits) t[7 * kThreads] += depth_cap_hits;
        if (p.verdict) p.verdict[idx] = (uint8_t)verdict;
        if (p.first_obstacle) p.first_obstacle[idx] = first;
        const double cost = RQCD_SLOT(SLOT_COST);
        if (verdict == 0 && (cost < bcost[tid] || (cost == bcost[tid] && bidx[tid] >= 0 && p.index_base + idx < bidx[tid]))) {  // lowest index on ties
          bcost[tid] = cost;
          bidx[tid] = p.index_base + idx;
        }
        have = false;
        pair_hits = 0;
        depth_cap_hits = 0;
      }
    }
  }
#undef RQCD_SLOT

  // ---- CTA partial
  __syncthreads();
  if (warp == 0) {
    double bc = __longlong_as_double(0x7ff0000000000000ll);
    long long bi = -1;
    for (int j = lane; j < kThreads; j += 32) {
      if (better(bcost[j], bidx[j], bc, bi)) {
        bc = bcost[j];
        bi = bidx[j];
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double oc = __shfl_down_sync(kFull, bc, off);
      const long long oi = __shfl_down_sync(kFull, bi, off);
      if (better(oc, oi, bc, bi)) {
        bc = oc;
        bi = oi;
      }
    }
    BlockPartial* out = p.partials + blockIdx.x;
    for (int j = 0; j < 8; j++) {
      unsigned long long s = 0;
      for (int q = lane; q < kThreads; q += 32) s += tcnt[j * kThreads + q];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(kFull, s, off);
      if (lane == 0) {
        if (j < 4) out->traj[j] = s;
        else out->pair[j - 4] = s;
      }
    }
    if (lane == 0) {
      out->best_cost = bc;
      out->best_index = bi;
    }
  }
}


// The CTA partials of a check launch -> the launch's summary. One block; counts by warp shuffles (shared-memory atomics on
// eight 64-bit words from 256 threads serialise: the first version of this kernel took 40 us for 296 partials), the best
// candidate by the same rule everywhere (lowest cost, lowest index on ties), partials visited in block order.
__global__ void __launch_bounds__(256) k_finalize(const BlockPartial* parts, int nparts, DevSummary* out) {
  __shared__ unsigned long long wsum[8][8];
  __shared__ double wcost[8];
  __shared__ long long widx[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double bc = __longlong_as_double(0x7ff0000000000000ll);
  long long bi = -1;
  unsigned long long loc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int j = tid; j < nparts; j += blockDim.x) {
#pragma unroll
    for (int v = 0; v < 4; v++) {
      loc[v] += parts[j].traj[v];
      loc[4 + v] += parts[j].pair[v];
    }
    if (better(parts[j].best_cost, parts[j].best_index, bc, bi)) {
      bc = parts[j].best_cost;
      bi = parts[j].best_index;
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
    for (int v = 0; v < 8; v++) loc[v] += __shfl_down_sync(kFull, loc[v], off);
    const double oc = __shfl_down_sync(kFull, bc, off);
    const long long oi = __shfl_down_sync(kFull, bi, off);
    if (better(oc, oi, bc, bi)) {
      bc = oc;
      bi = oi;
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int v = 0; v < 8; v++) wsum[warp][v] = loc[v];
    wcost[warp] = bc;
    widx[warp] = bi;
  }
  __syncthreads();
  if (tid == 0) {
    unsigned long long cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int w = 0; w < 8; w++) {
      for (int v = 0; v < 8; v++) cnt[v] += wsum[w][v];
      if (w > 0 && better(wcost[w], widx[w], bc, bi)) {
        bc = wcost[w];
        bi = widx[w];
      }
    }
    for (int v = 0; v < 4; v++) {
      out->traj[v] = cnt[v];
      out->pair[v] = cnt[4 + v];
    }
    out->pairs_checked = cnt[4] + cnt[5] + cnt[6] + cnt[7];
    out->best_cost = bc;
    out->best_index = bi;
  }
}

// ---- RQCD_F_FRAGILE: which results hinge on the last bits of the cubic solver's acos / cos / pow ---------------------
// Everything on the path is the reference's own IEEE operation sequence except those three calls (glibc there, polynomials
// here; two conforming libraries differ by an ulp or two). Where the closed-form roots are ill-conditioned -- grazing
// contact with coefficient ratios of 1e3 and more -- that ulp is amplified until it decides on which side of the tangent
// plane a computed critical point lands, and the reference's own answer is a coin toss. k_fragile finds those pairs.
//
// One thread per (trajectory i, obstacle k) pair, obstacle-major (the lanes of a warp read neighbouring candidates and one
// obstacle record). The thread walks the pair's recursion (end-point tests in the first section, depth first, low child
// parked: CollisionChecker.cpp:70-193) NINE times in lockstep: copy 0 as shipped, copies 1..8 with the three cosines of
// solveP3's trigonometric branch moved by +-eta in all eight sign patterns (the reference takes three independent cos()
// calls) and Cardano's cube root by the relative factor 1 +- eta. Every copy keeps its own interval ends -- a child
// interval is bounded by a computed root, so the copies' sections drift apart by what the perturbation did to that root.
//
// A section is a list of SIGN DECISIONS (SectionProbe::q): window against minTimeSection, midpoint inside the obstacle,
// quartic or cubic, the solver's own branches (SolverDecision), every root against ts / mid / tf, the plane side of every
// root and of both ends. Each comes with the magnitude of the terms it is the difference of (SectionProbe::s): its own
// rounding noise is a few ulps of that. The copies act as a numerical derivative of each quantity with respect to the
// transcendental outputs. A decision is PINNED when
//     |q(copy 0)| >= margin * max_v |q(copy v) - q(copy 0)| + floor_ulps * 2^-53 * s      (defaults: 4 and 8),
// i.e. when it is farther from zero than what the perturbation did to it AND than the rounding noise of its own
// evaluation (a root whose response to eta is below one ulp still moves by that ulp for some libm, and with it a
// plane-side value that sits on its noise floor -- the signature of grazing contact). Which decisions are NEEDED:
//   * those that cannot involve the transcendentals -- the first section's window, inside test, degree test and end-point
//     sides, and those of children whose interval ends are exact midpoints -- are exempt: they are the reference's bits;
//   * a root is a test point: where it lies relative to ts / mid / tf only matters if the root can be on the obstacle side
//     (its plane side is not pinned positive), or if the end it sits next to can (the child spawned by that end is bounded
//     by its neighbour in the list, :166, :183). A root of a child section often IS the child's start: on a flat face the
//     child's plane is the parent's;
//   * roots are kept by ORIGIN (two per quadratic factor of the quartic), not by sorted position, so a factor whose
//     discriminant changes sign under the perturbation -- an exact double root: every rest-to-rest primitive has one at
//     t = Tf, for every plane -- does not shift the others. Such a factor's two roots appear or vanish at its double-root
//     location -p/2; if the plane side THERE is pinned positive (and so are both ends'), test points that come and go
//     there decide nothing: the discriminant's sign is then not needed, nor are that factor's roots;
//   * the section's result and which children it has must be the same in all nine copies.
// A pair is MARKED at the first section where a needed decision is not pinned; unmarked pairs take copy 0's decisions --
// the shipped kernels' -- for every libm within eta.
enum ProbeSlot {
  PQ_WINDOW = 0,  // (tf - ts) - minT                              (:93-96)
  PQ_INSIDE,      // sphere: r - |Xm - c|; prism: min_i (h_i - |q_i|)   (:89-92)
  PQ_QUART,       // |pc[0]| - 1e-6                                (:47, :116)
  PQ_SOLVER,      // SD_COUNT entries
  PQ_ROOT = PQ_SOLVER + SD_COUNT,  // per root by origin: r - ts, r - mid, r - tf, plane side (4 x 4)
  PQ_DTS = PQ_ROOT + 16, PQ_DTF,   // plane side of the section's ends
  PQ_LOCA, PQ_LOCB,                // plane side at the double-root location of each quadratic factor
  PQ_COUNT
};
struct SectionProbe {
  int sig;          // StepResult | has_hi << 4 | has_lo << 5
  unsigned need;    // which q[] entries decide this section (before the rules above drop some)
  unsigned exists;  // bit j: root slot j holds a real root; bit 8: quartic (else cubic: slots 0..2)
  double t_next, lo_tf, mid, ts, tf;
  double loc[2];    // the double-root locations themselves
  double q[PQ_COUNT], s[PQ_COUNT];
};
// sum_d |n_d| (sum_k |c_kd| |t|^(5-k) + |point_d|): the magnitude of the terms of (X(t) - point) . n
RQCD_DEV double side_scale(const double (&c)[18], double t, V3 pp, V3 pn) {
  const double T = fabs(t);
  double ax = 0.0, ay = 0.0, az = 0.0;
#pragma unroll
  for (int k = 0; k < 6; k++) ax = ax * T + fabs(c[3 * k]), ay = ay * T + fabs(c[3 * k + 1]), az = az * T + fabs(c[3 * k + 2]);
  return fabs(pn.x) * (ax + fabs(pp.x)) + fabs(pn.y) * (ay + fabs(pp.y)) + fabs(pn.z) * (az + fabs(pp.z));
}
// `clean`: the section's ends do not depend on a computed root (the first section; children bounded by midpoints)
template <class Obs>
RQCD_DEV void probe_section(const double (&c)[18], double ts, double tf, const EndPoints ends, bool first, bool clean,
                            const Obs& obs, double minT, const Perturb pt, SectionProbe& o) {
  const double mid = (ts + tf) / 2;
  const V3 Xm = traj_value(c, mid);
  const V3 ctr = obs.center();
  const double rad = obs.radius();
  const bool prism = Obs::kSphereOnly ? false : obs.prism();
  const bool in_ends = first && ends_inside(obs, ends.xs(), ends.xf());
  __syncwarp();
  PlainMath pm;
  pm.pt = pt;
  pm.pt.diag = true;
  bool in_m;
  V3 pp, pn;
  double pc[5], r[4];
  const int n = plane_and_roots(pm, c, obs, Xm, ctr, rad, prism, in_m, pp, pn, pc, r[0], r[1], r[2], r[3]);
  const bool quart = fabs(pc[0]) > 1e-6;
  // roots by origin: the quartic's from the solver's diagnostics, the cubic's as returned (x0; x1, x2 when real)
  double ro[4];
  unsigned ex;
  if (quart) {
#pragma unroll
    for (int j = 0; j < 4; j++) ro[j] = pm.pt.ro[j];
    ex = (!(pm.pt.q[SD_DA] < 0.0) ? 3u : 0u) | (!(pm.pt.q[SD_DB] < 0.0) ? 12u : 0u) | 256u;
  } else {
    ro[0] = r[0], ro[1] = r[1], ro[2] = r[2], ro[3] = 0.0;
    ex = n >= 2 ? 7u : 1u;
  }
  Children ch;
  section_children(c, ts, tf, mid, pp, pn, pc, n, r[0], r[1], r[2], r[3], ends.amax(), !in_m, ch);
  int res = (ch.has_hi | ch.has_lo) ? STEP_CHILD : STEP_CLEAR;
  if (tf - ts < minT) res = STEP_INDETERM;
  if (in_m) res = STEP_COLLISION;
  if (in_ends) res = STEP_COLLISION;
  o.sig = res | ((res == STEP_CHILD) ? ((ch.has_hi ? 16 : 0) | (ch.has_lo ? 32 : 0)) : 0);
  o.t_next = ch.t_next, o.lo_tf = ch.lo_tf, o.mid = mid, o.ts = ts, o.tf = tf;
  o.loc[0] = pm.pt.loc[0], o.loc[1] = pm.pt.loc[1];
  o.exists = 0;
#pragma unroll
  for (int j = 0; j < PQ_COUNT; j++) o.q[j] = kNoDecision, o.s[j] = 0.0;
  o.q[PQ_WINDOW] = (tf - ts) - minT;
  o.s[PQ_WINDOW] = fabs(tf) + fabs(ts);
  const double xs = ((fabs(Xm.x) + fabs(Xm.y)) + fabs(Xm.z)) + ((fabs(ctr.x) + fabs(ctr.y)) + fabs(ctr.z));
  if (prism) {
    const V3 qm = rotate<Obs, true>(obs, Xm - ctr);
    o.q[PQ_INSIDE] = fmin(fmin(obs.half(0) - fabs(qm.x), obs.half(1) - fabs(qm.y)), obs.half(2) - fabs(qm.z));
    o.s[PQ_INSIDE] = 2.0 * xs + fmax(fmax(obs.half(0), obs.half(1)), obs.half(2));
  } else {
    o.q[PQ_INSIDE] = rad - norm2(Xm - ctr);
    o.s[PQ_INSIDE] = xs + rad;
  }
  unsigned need = 0;
  if (!in_ends) {  // a first section decided by the end points is the same in every copy
    if (!clean) need |= 1u << PQ_INSIDE;
    if (!in_m) {
      if (!clean) need |= 1u << PQ_WINDOW;
      if (res != STEP_INDETERM) {
        o.exists = ex;
        o.q[PQ_QUART] = fabs(pc[0]) - 1e-6;
        o.s[PQ_QUART] = 5.0 * ((fabs(pn.x * c[0]) + fabs(pn.y * c[1])) + fabs(pn.z * c[2]));
        // (SD_TRIG is reported but never needed: r^2 < q^3 is decided before any transcendental is taken, and at
        // r^2 = q^3 the two branches are continuations of each other -- a copy that takes the other one shows in the
        // scatter of everything downstream)
        if (!clean) need |= 1u << PQ_QUART;
        // (a quartic whose resolvent's single root dominates takes it whatever |x2| < eps says, see Perturb::x2_alt)
        if (!(quart && pm.pt.x2_alt > 0.0)) need |= 1u << (PQ_SOLVER + SD_X2);
        if (quart) need |= (((1u << SD_COUNT) - 1u) & ~((1u << SD_TRIG) | (1u << SD_X2))) << PQ_SOLVER;
#pragma unroll
        for (int j = 0; j < SD_COUNT; j++) o.q[PQ_SOLVER + j] = pm.pt.q[j], o.s[PQ_SOLVER + j] = pm.pt.qs[j];
        // plane side D(t) = (X(t) - point) . n (:157-158, :180-181) and what one ulp of t does to it: D'(t) is the velocity
        // polynomial the roots were computed from
#define RQCD_PROBE_SIDE(slot, t_)                                                                                \
  {                                                                                                             \
    const double t = (t_);                                                                                      \
    const double dv = fabs((((pc[0] * t + pc[1]) * t + pc[2]) * t + pc[3]) * t + pc[4]);                        \
    o.q[slot] = dot(traj_value(c, t) - pp, pn);                                                                 \
    o.s[slot] = side_scale(c, t, pp, pn) + dv * fabs(t);                                                        \
  }
#pragma unroll
        for (int i = 0; i < 4; i++) {
          if ((ex >> i) & 1u) {
            o.q[PQ_ROOT + 4 * i + 0] = ro[i] - ts, o.s[PQ_ROOT + 4 * i + 0] = fabs(ro[i]) + fabs(ts);
            o.q[PQ_ROOT + 4 * i + 1] = ro[i] - mid, o.s[PQ_ROOT + 4 * i + 1] = fabs(ro[i]) + fabs(mid);
            o.q[PQ_ROOT + 4 * i + 2] = ro[i] - tf, o.s[PQ_ROOT + 4 * i + 2] = fabs(ro[i]) + fabs(tf);
            RQCD_PROBE_SIDE(PQ_ROOT + 4 * i + 3, ro[i])
            need |= 7u << (PQ_ROOT + 4 * i);  // the caller drops comparisons that cannot matter
            if (ro[i] > ts && ro[i] < tf) need |= 8u << (PQ_ROOT + 4 * i);
          }
        }
        RQCD_PROBE_SIDE(PQ_DTS, ts)
        RQCD_PROBE_SIDE(PQ_DTF, tf)
        if (!clean) need |= (1u << PQ_DTS) | (1u << PQ_DTF);
        if (quart) {
          RQCD_PROBE_SIDE(PQ_LOCA, pm.pt.loc[0])
          RQCD_PROBE_SIDE(PQ_LOCB, pm.pt.loc[1])
        }
#undef RQCD_PROBE_SIDE
      }
    }
  }
  o.need = need;
}

template <int MODE, bool MOVING, bool PAIRWISE>
__global__ void __launch_bounds__(128) k_fragile(const CheckParams p, double eta, double margin, double floor_ulps, const int* first,
                                                 uint8_t* fragile, uint8_t* matrix, unsigned long long* why) {
  static_assert(PQ_COUNT <= 31, "the need mask is one word; reason 31 is the copies' disagreement");
  constexpr int kCopies = 9;
  const int lane = threadIdx.x & 31;
  const long long m = PAIRWISE ? 1 : p.m, total = p.n * m;
  const double* tab = reinterpret_cast<const double*>(p.obs_blob);
  const int* oflags = PAIRWISE ? nullptr : reinterpret_cast<const int*>(tab + (size_t)p.stride * p.mp);
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long j0 = blockIdx.x * (long long)blockDim.x + (threadIdx.x - lane); j0 < total; j0 += step) {
    const long long j = j0 + lane;
    bool alive = j < total;
    const long long i = alive ? j % p.n : 0;
    const int k = alive ? (int)(j / p.n) : 0;
    if (alive && p.skip && p.skip[i] != 0) alive = false;
    // a pair behind the trajectory's first hit cannot change the trajectory's result; it only matters to the matrix
    const bool counts = !first || first[i] < 0 || k <= first[i];
    if (!matrix && !counts) alive = false;
    double c[18], t0 = 0.0, t1 = 0.0, cost;
#pragma unroll
    for (int q = 0; q < 18; q++) c[q] = 0.0;
    if (alive) load_trajectory<MODE>(p, i, c, t0, t1, cost);
    LaneSphere sph = {0.0, 0.0, 0.0, 1.0};
    TableObs obs;
    obs.rec = PAIRWISE ? nullptr : tab + (size_t)k * p.stride;
    if (PAIRWISE && alive) {
      sph.cx = ld_in(p.spheres + i), sph.cy = ld_in(p.spheres + p.sphere_stride + i);
      sph.cz = ld_in(p.spheres + 2 * p.sphere_stride + i), sph.r = ld_in(p.spheres + 3 * p.sphere_stride + i);
    }
    if (MOVING && alive && (oflags[k] & OBS_MOVING)) {  // Trajectory::operator- (Trajectory.hpp:121-128)
#pragma unroll
      for (int q = 0; q < 18; q++) c[q] = c[q] - obs.rec[OF_MOTION + q];
      const double ms = obs.rec[OF_MT0], me = obs.rec[OF_MT1];
      t0 = (t0 < ms) ? ms : t0;
      t1 = (me < t1) ? me : t1;
      if (!(t0 <= t1)) alive = false;  // no common window: NoCollision without a section
    }
    double endrow[EndPoints::kRows];
    EndPoints ends;
    ends.slot = endrow, ends.stride = 1;
    ends.set_xs(traj_value(c, t0));
    ends.set_xf(traj_value(c, t1));
    ends.set_amax(side_filter_bound(c, t0, t1));
    double ts[kCopies], tf[kCopies];
    double2 stack[kCopies][kMaxDepth];
#pragma unroll
    for (int v = 0; v < kCopies; v++) ts[v] = t0, tf[v] = t1;
    int sp = 0;
    unsigned long long dirty_stack = 0;  // bit d: the interval parked at depth d has a root-derived end
    bool first_frame = true, frag = false, clean = true;
    while (__any_sync(kFull, alive)) {
      SectionProbe base;
      double spread[PQ_COUNT], t_next[kCopies], lo_tf[kCopies];
      unsigned ex_all = 0xffffffffu, ex_any = 0u;  // root slots that exist in every copy / in some copy
      bool differs = false;
#pragma unroll 1
      for (int v = 0; v < kCopies; v++) {
        Perturb pt;
        pt.on = v != 0;
        const int bits = v - 1;
        pt.d[0] = (bits & 1) ? eta : -eta, pt.d[1] = (bits & 2) ? eta : -eta, pt.d[2] = (bits & 4) ? eta : -eta;
        pt.a = 1.0 + pt.d[0];
        SectionProbe o;
        if (PAIRWISE) probe_section(c, ts[v], tf[v], ends, first_frame, clean, sph, p.min_t, pt, o);
        else probe_section(c, ts[v], tf[v], ends, first_frame, clean, obs, p.min_t, pt, o);
        ex_all &= o.exists, ex_any |= o.exists;
        if (v == 0) {
          base = o;
#pragma unroll
          for (int q = 0; q < PQ_COUNT; q++) spread[q] = 0.0;
        } else {
          differs = differs | (o.sig != base.sig);
#pragma unroll
          for (int q = 0; q < PQ_COUNT; q++) {
            // a root slot only compares where both copies have the root
            const bool root_slot = q >= PQ_ROOT && q < PQ_ROOT + 16;
            const bool both = !root_slot || (((o.exists & base.exists) >> ((q - PQ_ROOT) >> 2)) & 1u);
            const bool same = __double_as_longlong(o.q[q]) == __double_as_longlong(base.q[q]);
            const double d = fabs(o.q[q] - base.q[q]);
            // NaN differences (a NaN on one side only, inf - inf) count as unbounded
            if (both && !same) spread[q] = (d == d) ? fmax(spread[q], d) : __longlong_as_double(0x7ff0000000000000ll);
          }
        }
        t_next[v] = o.t_next, lo_tf[v] = o.lo_tf;
      }
      first_frame = false;
      if (alive) {
        int reason = differs ? 31 : -1;  // 31: the copies disagree on the section's result or its children
        unsigned pinned = 0;
#pragma unroll
        for (int q = 0; q < PQ_COUNT; q++) {
          // a NaN that every copy shares comes from the inputs, not from the transcendentals: every comparison with it is
          // false on every libm
          const bool shared_nan = base.q[q] != base.q[q] && spread[q] == 0.0;
          const bool ok = shared_nan || fabs(base.q[q]) >= margin * spread[q] + floor_ulps * 1.1102230246251565e-16 * base.s[q];
          pinned |= ok ? (1u << q) : 0u;
        }
        unsigned need = base.need;
        bool conflict = false;
        {
          // pinned sign of a plane-side value: +1 free side, -1 obstacle side, 0 not pinned (a clean section's ends are the
          // reference's own bits: pinned by construction)
          auto side = [&](int slot, bool exact) -> int {
            if (!(exact || ((pinned >> slot) & 1u))) return 0;
            return base.q[slot] > 0.0 ? 1 : (base.q[slot] <= 0.0 ? -1 : 0);
          };
          const int s_ts = side(PQ_DTS, clean), s_tf = side(PQ_DTF, clean);
          const bool quartic = (base.exists >> 8) & 1u;
          int s_root[4];
#pragma unroll
          for (int r_ = 0; r_ < 4; r_++) {
            const int b = PQ_ROOT + 4 * r_;
            s_root[r_] = ((base.exists >> r_) & 1u) ? side(b + 3, false) : 0;
            // where the root lies relative to an end does not matter when both are robustly on the same side of the plane
            // (whichever of the two is met first spawns the same child); relative to mid only when the root is free
            unsigned drop = 0;
            if (s_root[r_] != 0 && s_root[r_] == s_ts) drop |= 1u;
            if (s_root[r_] > 0) drop |= 2u;
            if (s_root[r_] != 0 && s_root[r_] == s_tf) drop |= 4u;
            need &= ~(drop << b);
          }
          if (quartic) {
            // a quadratic factor whose discriminant is not pinned: its two roots come and go at the double-root location
            const double width = base.tf - base.ts;
#pragma unroll
            for (int f = 0; f < 2; f++) {
              const int dslot = PQ_SOLVER + (f == 0 ? SD_DA : SD_DB), lslot = f == 0 ? PQ_LOCA : PQ_LOCB;
              const unsigned two = 3u << (2 * f);
              if (!((pinned >> dslot) & 1u)) {
                const double x = base.loc[f], near = 1e-5 * (width + fabs(x));
                const int s_loc = side(lslot, false);
                const bool own_ok = (!((base.exists >> (2 * f)) & 1u) || s_root[2 * f] == s_loc) &&
                                    (!((base.exists >> (2 * f + 1)) & 1u) || s_root[2 * f + 1] == s_loc);
                bool harmless = false;
                if (x < base.ts - near || x > base.tf + near) {
                  harmless = true;  // out of the section whether they exist or not
                } else if (fabs(x - base.tf) <= near) {
                  harmless = s_loc != 0 && s_loc == s_tf && own_ok;  // next to tf, on tf's side of the plane
                } else if (fabs(x - base.ts) <= near) {
                  harmless = s_loc != 0 && s_loc == s_ts && own_ok;
                } else if (fabs(x - base.mid) > near) {
                  // inside one half: free test points that come and go decide nothing as long as no point of that half's
                  // list can spawn a child bounded by them
                  const bool high = x > base.mid;
                  bool others_free = high ? (s_tf > 0) : (s_ts > 0);
#pragma unroll
                  for (int r_ = 0; r_ < 4; r_++) {
                    if (!((base.exists >> r_) & 1u) || (r_ >> 1) == f) continue;
                    const double rv = base.q[PQ_ROOT + 4 * r_ + 1];  // root - mid
                    const bool same_half = high ? !(rv < -near) : !(rv > near);
                    if (same_half && s_root[r_] <= 0) others_free = false;
                  }
                  harmless = s_loc > 0 && own_ok && others_free;
                }
                if (harmless) need &= ~((1u << dslot) | (0xffu << (PQ_ROOT + 8 * f)));
              } else if (((ex_all ^ ex_any) & two) != 0) {
                conflict = true;  // pinned discriminant, yet the copies disagree on the roots' existence
              }
            }
          } else if ((ex_all ^ ex_any) & 7u) {
            conflict = true;  // cubic: the copies disagree on the number of real roots
          }
        }
#pragma unroll
        for (int q = 0; q < PQ_COUNT; q++)
          if (((need >> q) & 1u) && !((pinned >> q) & 1u)) {
            differs = true;
            reason = reason < 0 ? q : reason;
          }
        if (conflict) {
          differs = true;
          reason = reason < 0 ? 31 : reason;
        }
        const int res = base.sig & 7;
        if (differs) {
          frag = true;
          alive = false;
          if (why) {
            const unsigned long long seen = atomicAdd(why + reason, 1ull);
#ifdef RQCD_FRAGILE_DEBUG
            if (seen < 4 && reason < PQ_COUNT)
              printf("fragile i %lld k %d reason %d q %.6g spread %.3g scale %.3g | ts %.17g tf %.17g sig %x sp %d clean %d ex %x exall %x exany %x | "
                     "Dts %.3g(%d) Dtf %.3g(%d) locA %.17g D %.3g sp %.3g s %.3g (%d) locB D %.3g (%d) | roots %.17g %.17g %.17g %.17g | Droot %.3g %.3g %.3g %.3g pinned %x\n",
                     i, k, reason, base.q[reason], spread[reason], base.s[reason], ts[0], tf[0], base.sig, sp, (int)clean, base.exists, ex_all, ex_any,
                     base.q[PQ_DTS], (int)((pinned >> PQ_DTS) & 1), base.q[PQ_DTF], (int)((pinned >> PQ_DTF) & 1), base.q[PQ_ROOT + 2] + tf[0],
                     base.q[PQ_LOCA], spread[PQ_LOCA], base.s[PQ_LOCA], (int)((pinned >> PQ_LOCA) & 1), base.q[PQ_LOCB], (int)((pinned >> PQ_LOCB) & 1),
                     base.q[PQ_ROOT + 2] + tf[0], base.q[PQ_ROOT + 6] + tf[0], base.q[PQ_ROOT + 10] + tf[0], base.q[PQ_ROOT + 14] + tf[0],
                     base.q[PQ_ROOT + 3], base.q[PQ_ROOT + 7], base.q[PQ_ROOT + 11], base.q[PQ_ROOT + 15], pinned);
#else
            (void)seen;
#endif
          }
        } else if (res == STEP_CHILD) {
          const bool hi = (base.sig & 16) != 0, lo = (base.sig & 32) != 0;
          if (hi & lo) {
            if (sp < kMaxDepth) {
#pragma unroll
              for (int v = 0; v < kCopies; v++) stack[v][sp] = make_double2(ts[v], lo_tf[v]);
              const bool lo_clean = clean && __double_as_longlong(base.lo_tf) == __double_as_longlong(base.mid);
              dirty_stack = (dirty_stack & ~(1ull << sp)) | (lo_clean ? 0ull : (1ull << sp));
              sp++;
            } else {
              alive = false;  // depth cap: reported by the check itself
            }
          }
          clean = clean && __double_as_longlong(base.t_next) == __double_as_longlong(base.mid);
#pragma unroll
          for (int v = 0; v < kCopies; v++) {
            if (hi) ts[v] = t_next[v];
            else tf[v] = t_next[v];
          }
        } else if (res == STEP_CLEAR && sp > 0) {
          sp--;
          clean = ((dirty_stack >> sp) & 1ull) == 0;
#pragma unroll
          for (int v = 0; v < kCopies; v++) ts[v] = stack[v][sp].x, tf[v] = stack[v][sp].y;
        } else {
          alive = false;  // the pair is decided, the same way in every copy
        }
      }
    }
    if (frag) {
      if (matrix) matrix[i * m + k] |= 0x80;
      if (fragile && counts) fragile[i] = 1;
    }
  }
}

// ---- multi-GPU exchange (rqcd_multi_*): a shard's summary goes to its slot of a table on the first device -- a store
// through peer-mapped memory (NVLink) or zero-copy host memory -- and one thread block there merges the table with
// rqcd_merge_summaries' rule: counts add; best = lowest cost, lowest index on ties (identical for every shard count).
__global__ void k_publish_summary(const DevSummary* mine, DevSummary* slot) {
  if (threadIdx.x == 0) {
    *slot = *mine;
    __threadfence_system();
  }
}
__global__ void k_merge_summaries(const DevSummary* slots, int n, DevSummary* out, DevSummary* copy) {
  if (threadIdx.x != 0) return;
  DevSummary m = {};
  m.best_cost = __longlong_as_double(0x7ff0000000000000ll);
  m.best_index = -1;
  for (int k = 0; k < n; k++) {
    const DevSummary s = slots[k];
    for (int v = 0; v < 4; v++) {
      m.traj[v] += s.traj[v];
      m.pair[v] += s.pair[v];
    }
    m.pairs_checked += s.pairs_checked;
    if (better(s.best_cost, s.best_index, m.best_cost, m.best_index)) {
      m.best_cost = s.best_cost;
      m.best_index = s.best_index;
    }
  }
  *out = m;
  if (copy) *copy = m;
}

// ---- RapidTrajectoryGenerator::Generate + GetCost + GetTrajectory only --------------------------------
// params (may be null): 9 rows alpha xyz, beta xyz, gamma xyz -- GetAxisParamAlpha/Beta/Gamma (.hpp:219-229), the values the
// generator computed, not back-computed from the coefficients
__global__ void __launch_bounds__(256) k_generate(const CheckParams p, double* cost, double* coeffs, long long coeff_stride,
                                                  double* params, long long param_stride) {
  const long long n = p.n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double c[18], t0, t1, co;
    load_trajectory<MODE_BC>(p, i, c, t0, t1, co);
    if (cost) cost[i] = co;
    if (coeffs) {
#pragma unroll
      for (int j = 0; j < 18; j++) coeffs[j * coeff_stride + i] = c[j];
    }
    if (params) {
      const BcIndex x = bc_index(p, i);
      const double Tf = bc_at(p, 18, x);
#pragma unroll
      for (int ax = 0; ax < 3; ax++) {
        const AxisParams q = gen_axis(bc_at(p, 0 + ax, x), bc_at(p, 3 + ax, x), bc_at(p, 6 + ax, x), bc_at(p, 9 + ax, x),
                                      bc_at(p, 12 + ax, x), bc_at(p, 15 + ax, x), (p.goal_mask >> (3 * ax)) & 1,
                                      (p.goal_mask >> (3 * ax + 1)) & 1, (p.goal_mask >> (3 * ax + 2)) & 1, Tf);
        params[(0 + ax) * param_stride + i] = q.a;
        params[(3 + ax) * param_stride + i] = q.b;
        params[(6 + ax) * param_stride + i] = q.g;
      }
    }
  }
}

// ---- CollisionChecker::CollisionCheck(Boundary) (src/CollisionChecker.cpp:25-68), documented intent: the
// trajectory is on the wrong side of the plane if (X(t) - point).normal <= 0 at t_start, t_end or at any time
// inside the window where its velocity along the normal vanishes. (The reference prepends the two end times to
// `roots` but then lets the solver overwrite them and reads two uninitialised slots, :43-55; the end times are
// tested here and nothing else is.) planes: prepared on the host, 6 doubles each: point, unit normal
// (Vec3::GetUnitVector's float-narrowed norm, :29).
template <int MODE>
__global__ void __launch_bounds__(256) k_planes(const CheckParams p, const double* planes, int np, uint8_t* verdict,
                                                uint8_t* matrix) {
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i0 = blockIdx.x * (long long)blockDim.x + (threadIdx.x & ~31); i0 < p.n; i0 += step) {  // warp-uniform
    const long long i = i0 + (threadIdx.x & 31);
    const bool live = i < p.n;
    const long long j = live ? i : p.n - 1;
    double c[18], t0, t1, cost;
    load_trajectory<MODE>(p, j, c, t0, t1, cost);
    const V3 Xs = traj_value(c, t0), Xf = traj_value(c, t1);
    int first_hit = 0;
    for (int k = 0; k < np; k++) {
      const double* q = planes + 6 * k;
      const V3 pp = v3(q[0], q[1], q[2]), pn = v3(q[3], q[4], q[5]);
      double r0, r1, r2, r3;
      const int n = critical_times(c, pn, r0, r1, r2, r3);
      bool hit = on_obstacle_side(Xs, pp, pn) | on_obstacle_side(Xf, pp, pn);
#define RQCD_PLANE_ROOT(idx, r) \
  if (n > idx && r >= t0 && r <= t1) hit = hit | on_obstacle_side(traj_value(c, r), pp, pn);
      RQCD_PLANE_ROOT(0, r0)
      RQCD_PLANE_ROOT(1, r1)
      RQCD_PLANE_ROOT(2, r2)
      RQCD_PLANE_ROOT(3, r3)
#undef RQCD_PLANE_ROOT
      __syncwarp();
      if (live && matrix) matrix[i * (long long)np + k] = hit ? 1 : 0;
      if (hit && !first_hit) first_hit = 1;
    }
    if (live && verdict) verdict[i] = (uint8_t)first_hit;
  }
}

// ---- RapidTrajectoryGenerator::CheckInputFeasibility (src/RapidTrajectoryGenerator.cpp:74-160) --------------------
RQCD_DEV void feas_axes(const CheckParams& p, long long i, AxisState (&ax)[3], double& Tf) {
  const BcIndex x = bc_index(p, i);
  Tf = bc_at(p, 18, x);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const double p0 = bc_at(p, 0 + k, x), v0 = bc_at(p, 3 + k, x), a0 = bc_at(p, 6 + k, x);
    const AxisParams q = gen_axis(p0, v0, a0, bc_at(p, 9 + k, x), bc_at(p, 12 + k, x), bc_at(p, 15 + k, x),
                                  (p.goal_mask >> (3 * k)) & 1, (p.goal_mask >> (3 * k + 1)) & 1,
                                  (p.goal_mask >> (3 * k + 2)) & 1, Tf);
    ax[k].a0 = a0;
    ax[k].a = q.a;
    ax[k].b = q.b;
    ax[k].g = q.g;
    acc_peak_times(ax[k], ax[k].pk0, ax[k].pk1);
  }
}

struct FeasArgs {
  double grav[3];
  double fmin, fmax, wmax, min_t;
};

// The reference caches each axis' acceleration peak times in the generator object and clears them only in
// Reset()/SetInitialState() (SingleAxisTrajectory.cpp:133-158): a driver that reuses ONE object for a batch
// (ForestPerformanceEval.cpp:177-209) evaluates every candidate of the batch with the peak times of the first
// candidate that reached GetMinMaxAcc on that axis. One thread per group start walks its group in order until all
// three axes are initialised (normally the first candidate) and writes the group's peak times to the group's ANCHOR
// slots: its first member and every kAnchorStep-th after it (anchor[q] = 1). A member finds its group's values by walking
// back to the nearest anchor (k_feas, at most kAnchorStep - 1 steps) -- 6 stores per group instead of 6 per member.
constexpr int kAnchorStep = 1024;
__global__ void __launch_bounds__(128) k_feas_group_peaks(const CheckParams p, const FeasArgs fa, const long long* group,
                                                          double* peaks /* 6 rows x n */, uint8_t* anchor /* n, zeroed */) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.n; i += (long long)gridDim.x * blockDim.x) {
    if (i > 0 && group[i] == group[i - 1]) continue;
    bool init[3] = {false, false, false};
    double pk[3][2] = {{0, 0}, {0, 0}, {0, 0}};
    long long j = i;
    for (; j < p.n && group[j] == group[i]; j++) {
      if (init[0] && init[1] && init[2]) continue;
      AxisState ax[3];
      double Tf;
      feas_axes(p, j, ax, Tf);
      double own[3][2];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        own[k][0] = ax[k].pk0;
        own[k][1] = ax[k].pk1;
        if (init[k]) {
          ax[k].pk0 = pk[k][0];
          ax[k].pk1 = pk[k][1];
        }
      }
      bool reached[3] = {false, false, false};
      feas_frame(ax, fa.grav, fa.fmin, fa.fmax, fa.wmax, 0.0, Tf, fa.min_t, reached);
#pragma unroll
      for (int k = 0; k < 3; k++)
        if (reached[k] && !init[k]) {
          init[k] = true;
          pk[k][0] = own[k][0];
          pk[k][1] = own[k][1];
        }
    }
    for (long long q = i; q < j; q += kAnchorStep) {
#pragma unroll
      for (int k = 0; k < 3; k++) {
        peaks[(2 * k) * p.n + q] = pk[k][0];
        peaks[(2 * k + 1) * p.n + q] = pk[k][1];
      }
      anchor[q] = 1;
    }
  }
}

// The section with its roots taken (feas_frame_sel), for the rare trips in which a lane's comparison is within the
// roundings of a limit: out of line, so that the trip loop stays small; the per-candidate values come back from the
// thread's shared-memory rows.
struct FeasAxes {
  double a0[3], a[3], b[3], g[3];
};
__device__ __noinline__ int feas_frame_literal(FeasAxes q, const double* hoist, FeasArgs fa, double t1, double t2) {
  AxisState ax[3];
  AxisHoist hz[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    ax[k].a0 = q.a0[k], ax[k].a = q.a[k], ax[k].b = q.b[k], ax[k].g = q.g[k];
    ax[k].pk0 = hoist[(6 * k + 0) * 128], ax[k].pk1 = hoist[(6 * k + 1) * 128];
    hz[k].ap0 = hoist[(6 * k + 2) * 128], hz[k].ap1 = hoist[(6 * k + 3) * 128];
    hz[k].tmax = hoist[(6 * k + 4) * 128], hz[k].jm = hoist[(6 * k + 5) * 128];
  }
  return feas_frame_sel(ax, hz, fa.grav, fa.fmin, fa.fmax, fa.wmax, t1, t2, fa.min_t);
}

// k_feas: CheckInputFeasibility for every candidate. Persistent warps, lane owns candidate: every trip of the warp loop
// runs ONE CheckInputFeasibilitySection per lane as one instruction stream (feas_frame_sel); a section that has to be
// split continues with its first half and parks the second on the lane's stack (the reference's recursion order: the
// second half is only looked at when the first is feasible, RapidTrajectoryGenerator.cpp:149-158). Candidates take 1 to
// some hundred sections (4.2 on average), so a lane whose candidate is decided takes the next one at once instead of
// waiting for the slowest of 32 -- from a per-warp QUEUE of ready-made candidates in shared memory: when the queue runs
// dry the whole warp builds 32 new ones, one per lane and converged (boundary conditions -> alpha, beta, gamma, peak
// times, and the values every section of the candidate would evaluate again: AxisHoist). Generation thus always runs
// with 32 busy lanes and the section trips with ~30.
// verdict / first (may be null): outputs of the check that follows, pre-filled for the candidates the filter rejects.
constexpr int kFeasRows = 32;  // per queued candidate: 3 x (a0, alpha, beta, gamma) + 3 x (pk0, pk1, AxisHoist) + Tf + index
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_feas(const CheckParams p, const FeasArgs fa, const double* peaks, uint8_t* result,
                                                    uint8_t* verdict, int* first, int refill, const uint8_t* anchor) {
  extern __shared__ __align__(16) double feas_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  double* hoist = feas_smem + threadIdx.x;                                  // [18][128]: this thread's candidate
  double* queue = feas_smem + 18 * 128 + warp * (kFeasRows * 32);          // [kFeasRows][32]: the warp's queue, slot-major rows
  AxisState ax[3];
#pragma unroll
  for (int k = 0; k < 3; k++) ax[k].a0 = ax[k].a = ax[k].b = ax[k].g = ax[k].pk0 = ax[k].pk1 = 0.0;
#pragma unroll
  for (int j = 0; j < 18; j++) hoist[j * 128] = 0.0;
  double2 stack[kMaxDepth];
  int sp = 0, q_n = 0;
  double t1 = 0.0, t2 = 0.0;
  long long idx = -1;
  bool have = false, exhausted = false;
  const FeasLimits lim = feas_limits(fa.fmin, fa.fmax, fa.wmax);
  for (;;) {
    unsigned idle = __ballot_sync(kFull, !have);
    if (__popc(idle) >= refill || idle == kFull) {
      for (int round = 0; round < 2; round++) {
        // ---- idle lanes pop from the top of the queue
        const int take_n = min(__popc(idle), q_n);
        const int rank = __popc(idle & lt_mask);
        if (!have && rank < take_n) {
          const double* e = queue + (q_n - 1 - rank);
#pragma unroll
          for (int k = 0; k < 3; k++) {
            ax[k].a0 = e[(4 * k + 0) * 32], ax[k].a = e[(4 * k + 1) * 32], ax[k].b = e[(4 * k + 2) * 32], ax[k].g = e[(4 * k + 3) * 32];
          }
#pragma unroll
          for (int j = 0; j < 18; j++) hoist[j * 128] = e[(12 + j) * 32];
          t1 = 0.0, t2 = e[30 * 32], idx = __double_as_longlong(e[31 * 32]), sp = 0, have = true;
        }
        q_n -= take_n;
        __syncwarp();
        idle = __ballot_sync(kFull, !have);
        if (idle == 0 || exhausted || round == 1) break;
        // ---- the queue is empty and lanes are still idle: the whole warp builds the next 32 candidates
        unsigned long long b = 0;
        if (lane == 0) b = atomicAdd(p.work_counter, 32ull);
        b = __shfl_sync(kFull, b, 0);
        if (b + 32ull >= (unsigned long long)p.n) exhausted = true;
        const long long my = (long long)b + lane;
        const int made = (b >= (unsigned long long)p.n) ? 0 : (int)min(32ll, p.n - (long long)b);
        {
          AxisState nx[3];
          double Tf = 0.0;
          feas_axes(p, my < p.n ? my : p.n - 1, nx, Tf);
          if (peaks && my < p.n) {  // the generator object's cached peak times: at the group's nearest anchor
            long long at = my;
            while (!anchor[at]) at--;
#pragma unroll
            for (int k = 0; k < 3; k++) nx[k].pk0 = peaks[(2 * k) * p.n + at], nx[k].pk1 = peaks[(2 * k + 1) * p.n + at];
          }
          double* e = queue + lane;  // slot = lane; slots [0, made) are valid
#pragma unroll
          for (int k = 0; k < 3; k++) {
            const AxisHoist hk = axis_hoist(nx[k]);
            e[(4 * k + 0) * 32] = nx[k].a0, e[(4 * k + 1) * 32] = nx[k].a, e[(4 * k + 2) * 32] = nx[k].b, e[(4 * k + 3) * 32] = nx[k].g;
            e[(12 + 6 * k + 0) * 32] = nx[k].pk0, e[(12 + 6 * k + 1) * 32] = nx[k].pk1;
            e[(12 + 6 * k + 2) * 32] = hk.ap0, e[(12 + 6 * k + 3) * 32] = hk.ap1;
            e[(12 + 6 * k + 4) * 32] = hk.tmax, e[(12 + 6 * k + 5) * 32] = hk.jm;
          }
          e[30 * 32] = Tf;
          e[31 * 32] = __longlong_as_double(my);
        }
        q_n = made;
        __syncwarp();
      }
    }
    if (__ballot_sync(kFull, have) == 0) {
      if (exhausted && q_n == 0) break;
      continue;
    }
    AxisHoist hz[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      ax[k].pk0 = hoist[(6 * k + 0) * 128], ax[k].pk1 = hoist[(6 * k + 1) * 128];
      hz[k].ap0 = hoist[(6 * k + 2) * 128], hz[k].ap1 = hoist[(6 * k + 3) * 128];
      hz[k].tmax = hoist[(6 * k + 4) * 128], hz[k].jm = hoist[(6 * k + 5) * 128];
    }
    // the section on squares (feas_frame_sq); the roots themselves only where a comparison is within the roundings
    bool open = false;
    int res = feas_frame_sq(ax, hz, fa.grav, lim, fa.fmax, fa.wmax, t1, t2, fa.min_t, open);
    if (!warp_all(!have | (lim.ok & !open))) {
      FeasAxes q;
#pragma unroll
      for (int k = 0; k < 3; k++) q.a0[k] = ax[k].a0, q.a[k] = ax[k].a, q.b[k] = ax[k].b, q.g[k] = ax[k].g;
      res = feas_frame_literal(q, hoist, fa, t1, t2);
    }
    if (have) {
      if (res == FEAS_SPLIT) {
        if (sp >= kMaxDepth) {
          res = 5;  // RQCD_INPUT_DEPTH_CAP
        } else {
          const double th = (t1 + t2) / 2;
          stack[sp++] = make_double2(th, t2);
          t2 = th;
        }
      } else if (res == RQCD_INPUT_FEASIBLE_ && sp > 0) {
        const double2 e = stack[--sp];
        t1 = e.x, t2 = e.y;
        res = FEAS_SPLIT;  // not done yet
      }
      if (res != FEAS_SPLIT) {
        result[idx] = (uint8_t)res;
        if (res != RQCD_INPUT_FEASIBLE_) {
          if (verdict) verdict[idx] = (uint8_t)255;  // RQCD_NOT_CHECKED
          if (first) first[idx] = -2;
        }
        have = false;
      }
    }
  }
}

// ---- the input filter's accept list: the indices i with result[i] == InputFeasible, ASCENDING (a stable compaction in
// three small launches: per-chunk counts, their exclusive scan, the writes). The check kernel takes its work from this
// list; ascending order keeps its reads of the boundary-condition rows as dense as the acceptance rate allows (a list
// appended in completion order cost 1.8x the DRAM traffic).
constexpr int kListChunk = 2048;  // candidates per block of 256 threads, 8 consecutive ones per thread
__global__ void __launch_bounds__(256) k_list_count(const uint8_t* result, long long n, unsigned* chunk_count) {
  const long long base = (long long)blockIdx.x * kListChunk + threadIdx.x * 8;
  unsigned c = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) c += (base + j < n && result[base + j] == RQCD_INPUT_FEASIBLE_) ? 1u : 0u;
  c = __reduce_add_sync(kFull, c);
  __shared__ unsigned ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) chunk_count[blockIdx.x] = ws[0] + ws[1] + ws[2] + ws[3] + ws[4] + ws[5] + ws[6] + ws[7];
}
// one block: chunk_count[0 .. nb) -> exclusive prefix sums in place (64-bit), total to *n_list
__global__ void __launch_bounds__(1024) k_list_scan(const unsigned* chunk_count, unsigned long long* chunk_base, int nb,
                                                    unsigned long long* n_list) {
  __shared__ unsigned long long ws[32];
  __shared__ unsigned long long carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0ull;
  __syncthreads();
  for (int b0 = 0; b0 < nb; b0 += 1024) {
    const int b = b0 + threadIdx.x;
    const unsigned long long v = b < nb ? chunk_count[b] : 0ull;
    unsigned long long x = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned long long y = __shfl_up_sync(kFull, x, off);
      if (lane >= off) x += y;
    }
    if (lane == 31) ws[warp] = x;
    __syncthreads();
    if (warp == 0) {
      unsigned long long w = ws[lane];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const unsigned long long y = __shfl_up_sync(kFull, w, off);
        if (lane >= off) w += y;
      }
      ws[lane] = w;
    }
    __syncthreads();
    const unsigned long long before = carry + (warp > 0 ? ws[warp - 1] : 0ull) + x - v;
    if (b < nb) chunk_base[b] = before;
    __syncthreads();
    if (threadIdx.x == 1023) carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_list = carry;
}
__global__ void __launch_bounds__(256) k_list_write(const uint8_t* result, long long n, const unsigned long long* chunk_base,
                                                    long long* list) {
  const long long base = (long long)blockIdx.x * kListChunk + threadIdx.x * 8;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned mask = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) mask |= (base + j < n && result[base + j] == RQCD_INPUT_FEASIBLE_) ? (1u << j) : 0u;
  const unsigned c = __popc(mask);
  unsigned x = c;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const unsigned y = __shfl_up_sync(kFull, x, off);
    if (lane >= off) x += y;
  }
  __shared__ unsigned ws[8];
  if (lane == 31) ws[warp] = x;
  __syncthreads();
  unsigned before = x - c;
  for (int w = 0; w < warp; w++) before += ws[w];
  long long* out = list + chunk_base[blockIdx.x] + before;
#pragma unroll
  for (int j = 0; j < 8; j++)
    if ((mask >> j) & 1u) *out++ = base + j;
}

// ---- planner glue: stage code per candidate (rqcd_plan_stage). A non-zero stage doubles as k_check's skip mask.
__global__ void __launch_bounds__(256) k_plan_mask(long long n, const double* cost, double max_cost, const uint8_t* feas,
                                                   const uint8_t* plane_hit, uint8_t* stage) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    uint8_t st = 0;
    if (!(cost[i] <= max_cost)) st = 1;            // cost pruned (a NaN cost is pruned too)
    else if (feas && feas[i] != 0) st = 2;         // input infeasible or indeterminable
    else if (plane_hit && plane_hit[i] != 0) st = 3;  // a plane boundary is violated
    stage[i] = st;
  }
}
__global__ void __launch_bounds__(256) k_plan_stage(long long n, const uint8_t* verdict, uint8_t* stage) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (stage[i] == 0 && verdict[i] != 0) stage[i] = (uint8_t)(3 + verdict[i]);  // 4 collision, 5 indeterminable, 6 depth cap
}

// ---- solver unit kernels (test surface of Quartic::solveP3 / solve_quartic) ---------------------------
__global__ void __launch_bounds__(256) k_solve(bool quartic, const double* coef, long long n, double* roots, int* count) {
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  // warp-uniform trip count: the solver reconverges with __syncwarp, so all 32 lanes stay in the loop and
  // lanes past the end work on a clamped index without storing
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i0 = blockIdx.x * (long long)blockDim.x + (threadIdx.x & ~31); i0 < n; i0 += step) {
    const long long i = i0 + (threadIdx.x & 31);
    const bool live = i < n;
    const long long j = live ? i : n - 1;
    if (quartic) {
      double r0 = nan, r1 = nan, r2 = nan, r3 = nan;
      const int cnt = solve_quartic(coef[j], coef[n + j], coef[2 * n + j], coef[3 * n + j], r0, r1, r2, r3);
      if (live) {
        count[i] = cnt;
        roots[i] = r0;
        roots[n + i] = r1;
        roots[2 * n + i] = cnt > 2 ? r2 : nan;
        roots[3 * n + i] = cnt > 2 ? r3 : nan;
        if (cnt == 0) roots[i] = roots[n + i] = nan;
      }
    } else {
      double x0 = nan, x1 = nan, x2 = nan;
      const int cnt = solve_p3(coef[j], coef[n + j], coef[2 * n + j], x0, x1, x2);
      if (live) {
        count[i] = cnt;
        roots[i] = x0;
        roots[n + i] = x1;
        roots[2 * n + i] = x2;
      }
    }
  }
}

// ---- synthetic boundary conditions: splitmix64 of (seed, row, global index); ranges of
// test/PerformanceEval.cpp:59-65 (pos0 = 0, everything else U(-4,4), Tf U(0.2,4)) ------------------------
RQCD_DEV unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__global__ void __launch_bounds__(256) k_synth_bc(unsigned long long seed, long long index_base, long long n,
                                                  double* bc, long long stride) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long g = (unsigned long long)(index_base + i);
#pragma unroll
    for (int r = 0; r < 19; r++) {
      const unsigned long long h = splitmix64(splitmix64(seed ^ ((unsigned long long)r * 0xD6E8FEB86659FD93ull)) + g);
      const double u = (double)(h >> 11) * 0x1.0p-53;
      double v;
      if (r < 3) v = 0.0;
      else if (r == 18) v = 0.2 + u * (4.0 - 0.2);
      else v = -4.0 + u * 8.0;
      bc[r * stride + i] = v;
    }
  }
}

// ---- FP64 issue ceilings: dependent-free DFMA chains, and DMUL+DADD chains (what -fmad=false code issues) --
template <bool FMA>
__global__ void __launch_bounds__(kPeakThreads) k_fp64_peak(int iters, double* sink) {
  double a[kPeakChains];
  const double x = 1.0 + 1e-9 * threadIdx.x, y = 1e-9;
#pragma unroll
  for (int j = 0; j < kPeakChains; j++) a[j] = 1.0 + j;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < kPeakInnerOps; u++) {
#pragma unroll
      for (int j = 0; j < kPeakChains; j++) {
        if (FMA) a[j] = __fma_rn(a[j], x, y);
        else a[j] = (u & 1) ? __dadd_rn(a[j], y) : __dmul_rn(a[j], x);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < kPeakChains; j++) s += a[j];
  if (s == 123.456) sink[0] = s;
}

__device__ double g_rcp_const[3];
__global__ void k_init_constants() {
  g_rcp_const[0] = rcp_refined(3.0);
  g_rcp_const[1] = rcp_refined(9.0);
  g_rcp_const[2] = rcp_refined(54.0);
}

template <int MODE, bool MOVING, bool PAIRWISE, bool ENDBITS>
const void* check_fn() {
  return reinterpret_cast<const void*>(&k_check<MODE, MOVING, PAIRWISE, ENDBITS>);
}

// The end-point pre-pass pays off when a trajectory meets many obstacles; a handful (Forest: 5, early exit) is
// cheaper inside the pairs' first frames.
bool use_endbits(bool moving, bool pairwise, int m) { return !moving && !pairwise && m >= kEndBitsMinObstacles; }

const void* select_check(int mode, bool moving, bool pairwise, int m) {
  if (pairwise) return check_fn<MODE_BC, false, true, false>();
  if (moving) return mode == MODE_BC ? check_fn<MODE_BC, true, false, false>() : check_fn<MODE_COEFF, true, false, false>();
  if (use_endbits(moving, pairwise, m))
    return mode == MODE_BC ? check_fn<MODE_BC, false, false, true>() : check_fn<MODE_COEFF, false, false, true>();
  return mode == MODE_BC ? check_fn<MODE_BC, false, false, false>() : check_fn<MODE_COEFF, false, false, false>();
}

// SMs of the current device (the caller's DeviceGuard has selected the context's GPU): grid caps are multiples of it
int sm_count() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      sms < 1)
    sms = 148;  // B200
  return sms;
}

int grid_for(long long n, int cap, int threads = 256) {
  long long b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return (int)b;
}

}  // namespace

cudaError_t init_device_constants(cudaStream_t s) {
  k_init_constants<<<1, 1, 0, s>>>();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  void* src = nullptr;
  e = cudaGetSymbolAddress(&src, g_rcp_const);
  if (e != cudaSuccess) return e;
  e = cudaMemcpyToSymbolAsync(kRcpConst, src, sizeof(double) * 3, 0, cudaMemcpyDeviceToDevice, s);
  if (e != cudaSuccess) return e;
  return cudaStreamSynchronize(s);
}

size_t check_smem_bytes(int mp, int stride, bool moving) {
  size_t b = (size_t)stride * mp * 8 + (size_t)mp * 4;  // table + flags (mp % 4 == 0 keeps 16-byte alignment)
  b += (size_t)8 * kThreads * 8;                          // verdict counters, one row of 8 per thread
  b += (size_t)kThreads * 16;                             // best cost / index per thread
  b += (size_t)kThreads * 8 * EndPoints::kRows;           // X(t_start), X(t_end), side-filter bound per thread
  if (moving) b += (size_t)18 * kThreads * 8;             // the lane's own coefficients
  else b += (size_t)((mp + 31) / 32) * kThreads * 4;      // end-point test bits, one per obstacle and thread
  return b;
}

cudaError_t check_prepare(int mode, bool moving, bool pairwise, int m, size_t smem) {
  return cudaFuncSetAttribute(select_check(mode, moving, pairwise, m), cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)smem);
}

int check_max_blocks_per_sm(int mode, bool moving, bool pairwise, int m, size_t smem) {
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, select_check(mode, moving, pairwise, m), kThreads, smem) !=
      cudaSuccess)
    return 0;
  return nb;
}

const char* check_kernel_name(bool moving, bool pairwise, int m) {
  if (pairwise) return "k_check<pairwise>";
  if (moving) return "k_check<moving>";
  return use_endbits(moving, pairwise, m) ? "k_check<static,endbits>" : "k_check<static>";
}

cudaError_t launch_check(const CheckParams& p, int mode, bool moving, bool pairwise, int grid, cudaStream_t s) {
  const size_t smem = check_smem_bytes(pairwise ? 0 : p.mp, pairwise ? 0 : p.stride, moving);
  CheckParams q = p;
  if (pairwise) {
    q.mp = 0;
    q.stride = 0;
  }
  void* args[] = {&q};
  return cudaLaunchKernel(select_check(mode, moving, pairwise, p.m), dim3(grid), dim3(kThreads), args, smem, s);
}

cudaError_t launch_publish_summary(const DevSummary* mine, DevSummary* slot, cudaStream_t s) {
  k_publish_summary<<<1, 32, 0, s>>>(mine, slot);
  return cudaGetLastError();
}
cudaError_t launch_merge_summaries(const DevSummary* slots, int n, DevSummary* out, DevSummary* copy, cudaStream_t s) {
  k_merge_summaries<<<1, 32, 0, s>>>(slots, n, out, copy);
  return cudaGetLastError();
}

cudaError_t launch_fragile(const CheckParams& p, int mode, bool moving, bool pairwise, double eta, double margin, double floor_ulps,
                           const int* first, uint8_t* fragile, uint8_t* matrix, unsigned long long* why, cudaStream_t s) {
  const long long total = p.n * (long long)(pairwise ? 1 : p.m);
  if (total <= 0) return cudaSuccess;
  const int grid = grid_for(total, sm_count() * 16, 128);
  const bool bc = mode == MODE_BC;
#define RQCD_FRAGILE(MODE, MOVING, PAIRWISE) k_fragile<MODE, MOVING, PAIRWISE><<<grid, 128, 0, s>>>(p, eta, margin, floor_ulps, first, fragile, matrix, why)
  if (pairwise) {
    if (bc) RQCD_FRAGILE(MODE_BC, false, true);
    else RQCD_FRAGILE(MODE_COEFF, false, true);
  } else if (moving) {
    if (bc) RQCD_FRAGILE(MODE_BC, true, false);
    else RQCD_FRAGILE(MODE_COEFF, true, false);
  } else {
    if (bc) RQCD_FRAGILE(MODE_BC, false, false);
    else RQCD_FRAGILE(MODE_COEFF, false, false);
  }
#undef RQCD_FRAGILE
  return cudaGetLastError();
}

cudaError_t launch_finalize(const BlockPartial* partials, int grid, DevSummary* d_summary, cudaStream_t s) {
  k_finalize<<<1, 256, 0, s>>>(partials, grid, d_summary);
  return cudaGetLastError();
}

cudaError_t launch_planes(const CheckParams& p, int mode, const double* planes, int np, uint8_t* verdict, uint8_t* matrix,
                          cudaStream_t s) {
  if (p.n <= 0) return cudaSuccess;
  const int grid = grid_for(p.n, sm_count() * 8);
  if (mode == MODE_BC) k_planes<MODE_BC><<<grid, 256, 0, s>>>(p, planes, np, verdict, matrix);
  else k_planes<MODE_COEFF><<<grid, 256, 0, s>>>(p, planes, np, verdict, matrix);
  return cudaGetLastError();
}

cudaError_t launch_feasibility(const CheckParams& p, const double grav[3], double fmin, double fmax, double wmax,
                               double min_t, const long long* group, double* peaks, uint8_t* result, uint8_t* verdict, int* first,
                               long long* list, unsigned long long* n_list, cudaStream_t s) {
  if (p.n <= 0) return cudaSuccess;
  FeasArgs fa;
  for (int k = 0; k < 3; k++) fa.grav[k] = grav[k];
  fa.fmin = fmin, fa.fmax = fmax, fa.wmax = wmax, fa.min_t = min_t;
  const int grid = (int)((p.n + 127) / 128 < sm_count() * 16 ? (p.n + 127) / 128 : sm_count() * 16);
  uint8_t* anchor = group ? reinterpret_cast<uint8_t*>(peaks + 6 * p.n) : nullptr;  // the caller's buffer holds 6 n doubles + n bytes
  if (group) {
    cudaError_t e = cudaMemsetAsync(anchor, 0, (size_t)p.n, s);
    if (e != cudaSuccess) return e;
    k_feas_group_peaks<<<grid, 128, 0, s>>>(p, fa, group, peaks, anchor);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  cudaError_t e = cudaMemsetAsync(p.work_counter, 0, sizeof(unsigned long long), s);
  if (e != cudaSuccess) return e;
  static int refill = [] {
    const char* v = getenv("RQCD_FEAS_REFILL");
    return (v && atoi(v) >= 1 && atoi(v) <= 32) ? atoi(v) : 2;
  }();
  static int blocks_per_sm = [] {
    const char* v = getenv("RQCD_FEAS_BLOCKS");
    return (v && atoi(v) >= 1 && atoi(v) <= 16) ? atoi(v) : 8;
  }();
  const long long want = (p.n + 127) / 128;
  const int pgrid = (int)(want < (long long)sm_count() * blocks_per_sm ? want : (long long)sm_count() * blocks_per_sm);
  static int occ = [] {
    const char* v = getenv("RQCD_FEAS_OCC");
    return (v && atoi(v) >= 4 && atoi(v) <= 6) ? atoi(v) : 4;
  }();
  const size_t smem = (size_t)(18 * 128 + 4 * kFeasRows * 32) * sizeof(double);
  const void* fn = occ == 6 ? (const void*)&k_feas<6> : (occ == 5 ? (const void*)&k_feas<5> : (const void*)&k_feas<4>);
  e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const double* pk = group ? peaks : nullptr;
  void* args[] = {(void*)&p, (void*)&fa, (void*)&pk, (void*)&result, (void*)&verdict, (void*)&first, (void*)&refill, (void*)&anchor};
  e = cudaLaunchKernel(fn, dim3(pgrid), dim3(128), args, smem, s);
  if (e != cudaSuccess || !list) return e;
  // the accept list: n_list[0] = count, n_list[1 ..] = scratch (chunk bases, then chunk counts; see feasibility_list_words)
  const int nb = (int)((p.n + kListChunk - 1) / kListChunk);
  unsigned long long* chunk_base = n_list + 2;
  unsigned* chunk_count = reinterpret_cast<unsigned*>(chunk_base + nb);
  k_list_count<<<nb, 256, 0, s>>>(result, p.n, chunk_count);
  k_list_scan<<<1, 1024, 0, s>>>(chunk_count, chunk_base, nb, n_list);
  k_list_write<<<nb, 256, 0, s>>>(result, p.n, chunk_base, list);
  return cudaGetLastError();
}

// 64-bit words of scratch launch_feasibility needs behind n_list for n candidates
size_t feasibility_list_words(long long n) {
  const size_t nb = (size_t)((n + kListChunk - 1) / kListChunk);
  return 2 + nb + (nb + 1) / 2;
}

cudaError_t launch_plan_mask(long long n, const double* cost, double max_cost, const uint8_t* feas, const uint8_t* plane_hit,
                             uint8_t* stage, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  k_plan_mask<<<grid_for(n, sm_count() * 8), 256, 0, s>>>(n, cost, max_cost, feas, plane_hit, stage);
  return cudaGetLastError();
}
cudaError_t launch_plan_stage(long long n, const uint8_t* verdict, uint8_t* stage, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  k_plan_stage<<<grid_for(n, sm_count() * 8), 256, 0, s>>>(n, verdict, stage);
  return cudaGetLastError();
}

cudaError_t launch_generate(const CheckParams& p, double* cost, double* coeffs, long long coeff_stride, double* params,
                            long long param_stride, cudaStream_t s) {
  if (p.n <= 0) return cudaSuccess;
  k_generate<<<grid_for(p.n, sm_count() * 8), 256, 0, s>>>(p, cost, coeffs, coeff_stride, params, param_stride);
  return cudaGetLastError();
}

cudaError_t launch_solve(bool quartic, const double* coef, long long n, double* roots, int* count, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  k_solve<<<grid_for(n, sm_count() * 8), 256, 0, s>>>(quartic, coef, n, roots, count);
  return cudaGetLastError();
}

// rqcd_solve_conditioning: the solver of k_solve nine times per problem (see k_fragile): copy 0 as shipped, eight with the
// transcendental outputs moved by +-eta. cond = the largest movement of any reported root divided by eta: how many units a
// root moves per unit of libm error. flag = a branch decision of the solver (SolverDecision) is not pinned, or the
// copies disagree on the root count: the reference's own count is a coin toss there.
__global__ void __launch_bounds__(128) k_solve_cond(bool quartic, const double* coef, long long n, double eta, double margin,
                                                    double floor_ulps, double* cond, uint8_t* flag) {
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i0 = blockIdx.x * (long long)blockDim.x + (threadIdx.x & ~31); i0 < n; i0 += step) {
    const long long i = i0 + (threadIdx.x & 31);
    const bool live = i < n;
    const long long j = live ? i : n - 1;
    const double a = coef[j], b = coef[n + j], c = coef[2 * n + j], d = quartic ? coef[3 * n + j] : 0.0;
    double base_r[4] = {0, 0, 0, 0}, base_q[SD_COUNT], base_s[SD_COUNT], spread[SD_COUNT], moved = 0.0;
    int base_cnt = 0;
    bool bad = false;
#pragma unroll 1
    for (int v = 0; v < 9; v++) {
      PlainMath pm;
      pm.pt.on = v != 0;
      pm.pt.diag = true;
      const int bits = v - 1;
      pm.pt.d[0] = (bits & 1) ? eta : -eta, pm.pt.d[1] = (bits & 2) ? eta : -eta, pm.pt.d[2] = (bits & 4) ? eta : -eta;
      pm.pt.a = 1.0 + pm.pt.d[0];
      double r[4] = {0, 0, 0, 0};
      const int cnt = quartic ? solve_quartic(pm, a, b, c, d, r[0], r[1], r[2], r[3]) : solve_p3(pm, a, b, c, r[0], r[1], r[2]);
      const int used = quartic ? cnt : 3;  // the cubic reports Re / Im of the complex pair in x[1], x[2]
      if (v == 0) {
        base_cnt = cnt;
#pragma unroll
        for (int q = 0; q < 4; q++) base_r[q] = r[q];
#pragma unroll
        for (int q = 0; q < SD_COUNT; q++) base_q[q] = pm.pt.q[q], base_s[q] = pm.pt.qs[q], spread[q] = 0.0;
      } else {
        bad = bad | (cnt != base_cnt);
#pragma unroll
        for (int q = 0; q < 4; q++)
          if (q < used && cnt == base_cnt && __double_as_longlong(r[q]) != __double_as_longlong(base_r[q])) {
            const double dr = fabs(r[q] - base_r[q]);
            moved = (dr == dr) ? fmax(moved, dr) : __longlong_as_double(0x7ff0000000000000ll);
          }
#pragma unroll
        for (int q = 0; q < SD_COUNT; q++)
          if (__double_as_longlong(pm.pt.q[q]) != __double_as_longlong(base_q[q])) {
            const double dq = fabs(pm.pt.q[q] - base_q[q]);
            spread[q] = (dq == dq) ? fmax(spread[q], dq) : __longlong_as_double(0x7ff0000000000000ll);
          }
      }
    }
#pragma unroll
    for (int q = 0; q < SD_COUNT; q++) {
      if (q == SD_TRIG) continue;  // decided before any transcendental: the reference's own bits
      if (!quartic && q != SD_X2) continue;
      const bool shared_nan = base_q[q] != base_q[q] && spread[q] == 0.0;
      if (!shared_nan && !(fabs(base_q[q]) >= margin * spread[q] + floor_ulps * 1.1102230246251565e-16 * base_s[q])) bad = true;
    }
    if (live) {
      cond[i] = moved / eta;
      flag[i] = bad ? 1 : 0;
    }
  }
}

cudaError_t launch_solve_cond(bool quartic, const double* coef, long long n, double eta, double margin, double floor_ulps,
                              double* cond, uint8_t* flag, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  k_solve_cond<<<grid_for(n, sm_count() * 16, 128), 128, 0, s>>>(quartic, coef, n, eta, margin, floor_ulps, cond, flag);
  return cudaGetLastError();
}

cudaError_t launch_synth_bc(unsigned long long seed, long long index_base, long long n, double* bc, long long stride,
                            cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  k_synth_bc<<<grid_for(n, sm_count() * 8), 256, 0, s>>>(seed, index_base, n, bc, stride);
  return cudaGetLastError();
}

cudaError_t launch_fp64_peak(bool fma, int blocks, int iters, double* sink, cudaStream_t s) {
  if (fma) k_fp64_peak<true><<<blocks, kPeakThreads, 0, s>>>(iters, sink);
  else k_fp64_peak<false><<<blocks, kPeakThreads, 0, s>>>(iters, sink);
  return cudaGetLastError();
}

}  // namespace rqcd
