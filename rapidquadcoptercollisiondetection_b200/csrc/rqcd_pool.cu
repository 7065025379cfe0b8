// rqcd_pool.cu -- k_pool, the staged check kernel for obstacle TABLES (static or moving), sm_100a. Compile with
// -fmad=false (see rqcd_device.cuh). Same verdicts as k_check (rqcd_kernels.cu), which stays the kernel of the
// pairwise shape and of tables that do not fit next to the trajectory rows.
//
// The reference's CollisionCheckSection (src/CollisionChecker.cpp:82-193) is cut into two STAGES:
//   PLANE stage   X(mid), IsPointInside(mid), the tf - ts < minT test, GetTangentPlane(mid) (:87-99) and the certified
//                 solver skip (bernstein_clear, rqcd_device.cuh): when the plane-side quintic is positive on the whole
//                 interval, no test point of the reference can be on the obstacle side, whatever the roots are, and the
//                 section is NoCollision without children. 80 % of first sections and 50 % of deeper ones end here.
//   SOLVE stage   the rest (:103-191): critical times (quartic / cubic), sort, classification, plane-side tests, children.
// A warp works on 32 candidates at a time (its batch; their coefficients sit in per-thread shared-memory rows so that
// any lane can work for any trajectory) and chooses one of three kinds of TRIP, each one converged instruction stream:
//   A trip   every lane runs the PLANE stage of the first section of (its own trajectory, obstacle k), k the same for the
//            whole warp: obstacle record = broadcast loads, X(mid) of the whole window read from the row (evaluated once
//            per trajectory). Pairs that need the solver become TASKS in the warp's pool: (owner lane, obstacle, plane).
//   S trip   lanes holding a task in state NEED_SOLVE run the SOLVE stage; a section with children continues on the same
//            lane in state NEED_PLANE (depth first, the low child parked on the lane's private stack -- the reference's
//            recursion order); idle lanes first take tasks from the pool.
//   P trip   lanes in state NEED_PLANE run the PLANE stage of their child section.
// A trip of one kind is taken when enough lanes wait for it (CheckParams::th_solve / th_plane), A trips fill the gaps,
// so that every trip runs with most of its lanes busy although only 1 section in 4 needs the solver.
// A verdict other than NoCollision goes to the owner with one shared-memory atomicMin on (obstacle << 2 | verdict): the
// trajectory's verdict is the one of the lowest obstacle index (test/ForestPerformanceEval.cpp:216-225). With
// RQCD_F_EARLY_EXIT tasks and first sections of obstacles behind a known hit are dropped, and the executed-pair counts
// are the driver's (pairs up to and including the first hit).
// Instruction cache: the trip loop is ~30 KB of SASS against a 32 KB cache, and four warps per scheduler sit on different
// trips. Everything that is not needed by the common call is kept out of it: the PLAIN variant compiles the early-exit
// tests and verdict-matrix stores out, the rare side paths of a section are calls (RQCD_OUTLINE_COLD), the compiler's own
// division / sqrt fallbacks are out-of-line functions. Each of these was worth 5-11 %.
#define RQCD_OUTLINE_COLD  // the rare side paths of a section as calls, see rqcd_device.cuh
#include "rqcd_kernel_util.cuh"

namespace rqcd {

namespace {

constexpr int kPoolCap = 64;  // tasks per warp; an A trip adds at most 32 and is only taken when 32 slots are free
// per-thread shared-memory rows
enum PoolRow {
  PR_XM = 0,    // X(mid of the whole window): the first section's midpoint, the same for every obstacle (static tables)
  PR_T0 = 3, PR_T1 = 4,
  PR_AMAX = 5,  // side_filter_bound of the trajectory (static tables)
  PR_C = 6,     // 18 coefficients
  PR_PL = 24,   // plane (point, normal) of the task this LANE is running, between its PLANE and SOLVE stages
  kPoolRows = 30
};
constexpr int kQueueCap = 64;  // first sections waiting for their PLANE stage, per warp (owner << 16 | obstacle)
enum LaneState { LS_IDLE = 0, LS_NEED_PLANE = 1, LS_NEED_SOLVE = 2 };
enum PlaneStage { PS_CLEAR = 0, PS_COLLISION = 1, PS_INDETERM = 2, PS_SOLVE = 4 };

// The compiler's own division and square root for a whole section when an operand of any lane left the range of the
// regrouped sequences (see FastMath): rare, and kept out of line so that the hot loop stays inside the instruction cache.
// (Arguments and results travel by value, the coefficients are re-read from the owner's rows: nothing of the caller's
// register state needs an address.)
struct PlainPlane {
  V3 pp, pn;
  bool in_m;
};
__device__ __noinline__ PlainPlane tangent_plane_plain(const double* rec, V3 Xm) {
  TableObs obs;
  obs.rec = rec;
  PlainMath pm;
  PlainPlane o;
  tangent_plane(pm, obs, Xm, obs.center(), obs.radius(), obs.prism(), o.in_m, o.pp, o.pn);
  return o;
}
struct PlainRoots {
  double pc[5], r[4];
  int n;
};
__device__ __noinline__ PlainRoots critical_times_plain(const double* crow, const double* motion, int row_stride, V3 pn) {
  double c[18];
#pragma unroll
  for (int j = 0; j < 18; j++) c[j] = crow[(PR_C + j) * row_stride];
  if (motion) {
#pragma unroll
    for (int j = 0; j < 18; j++) c[j] = c[j] - motion[j];
  }
  PlainMath pm;
  PlainRoots o;
  o.n = critical_times(pm, c, pn, o.pc, o.r[0], o.r[1], o.r[2], o.r[3]);
  return o;
}

// One axis of RapidTrajectoryGenerator::Generate + GetTrajectory (gen_axis + axis_coeffs), out of line: the eight
// goal-flag branches exist once instead of three times in the batch-build code, which shares the instruction cache
// with the trip loop. Same arithmetic as load_trajectory (rqcd_kernel_util.cuh).
struct AxisOut {
  double k[6], cost;
};
__device__ __noinline__ AxisOut gen_axis_out(double p0, double v0, double a0, double pf, double vf, double af, unsigned goal3,
                                             double Tf) {
  const AxisParams q = gen_axis(p0, v0, a0, pf, vf, af, goal3 & 1u, (goal3 >> 1) & 1u, (goal3 >> 2) & 1u, Tf);
  AxisOut o;
  axis_coeffs(q, p0, v0, a0, o.k[0], o.k[1], o.k[2], o.k[3], o.k[4], o.k[5]);
  o.cost = q.cost;
  return o;
}
template <int MODE>
RQCD_DEV void load_trajectory_compact(const CheckParams& p, long long i, double (&c)[18], double& t0, double& t1, double& cost) {
  if (MODE == MODE_BC) {
    const BcIndex x = bc_index(p, i);
    const double Tf = bc_at(p, 18, x);
    double total = 0.0;
#pragma unroll
    for (int ax = 0; ax < 3; ax++) {
      const AxisOut q = gen_axis_out(bc_at(p, 0 + ax, x), bc_at(p, 3 + ax, x), bc_at(p, 6 + ax, x), bc_at(p, 9 + ax, x),
                                     bc_at(p, 12 + ax, x), bc_at(p, 15 + ax, x), (p.goal_mask >> (3 * ax)) & 7u, Tf);
#pragma unroll
      for (int j = 0; j < 6; j++) c[3 * j + ax] = q.k[j];
      total = (ax == 0) ? q.cost : total + q.cost;  // GetCost: cost[0] + cost[1] + cost[2] (.hpp:214-216)
    }
    cost = total;
    t0 = 0.0;
    t1 = Tf;
  } else {
    load_trajectory<MODE>(p, i, c, t0, t1, cost);
  }
}

// The PLANE stage of one section (CollisionChecker.cpp:89-99 + the certified skip). All 32 lanes call it together.
RQCD_DEV int plane_stage(const double (&c)[18], const TableObs& obs, V3 Xm, double ts, double tf, double amax, double minT,
                         bool active, V3& pp, V3& pn) {
  const V3 ctr = obs.center();
  const double rad = obs.radius();
  const bool prism = obs.prism();
  bool in_m;
  {
    FastMath f;
    tangent_plane(f, obs, Xm, ctr, rad, prism, in_m, pp, pn);
    if (!warp_all(f.ok | !active)) {
      const PlainPlane q = tangent_plane_plain(obs.rec, Xm);
      in_m = q.in_m, pp = q.pp, pn = q.pn;
    }
  }
  const bool clear = bernstein_clear(c, pp, pn, ts, tf, amax);
  int r = clear ? PS_CLEAR : PS_SOLVE;
  r = (tf - ts < minT) ? PS_INDETERM : r;  // :93-96
  r = in_m ? PS_COLLISION : r;             // :89-92
  return r;
}

// PLAIN: all pairs, no verdict matrix -- the common call; its early-exit tests and matrix stores are compiled out (the trip
// loop does not fit the instruction cache, every branch it does not contain shows).
template <int MODE, bool MOVING, int T, bool PLAIN>
__global__ void __launch_bounds__(T, 512 / T) k_pool(const CheckParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int kWarps = T / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wbase = tid - lane;
  const int m = p.m, mp = p.mp;
  // Shared-memory layout: everything whose size only depends on T first, at compile-time offsets (a pointer that has to
  // be recomputed from p.stride and p.mp at every use costs a dozen instructions under this kernel's register
  // pressure), then the table, its flags and the end-point bits.
  unsigned long long* wcnt = reinterpret_cast<unsigned long long*>(smem_raw);         // [kWarps][8]
  double* bcost = reinterpret_cast<double*>(wcnt + kWarps * 8);                       // [T]
  long long* bidx = reinterpret_cast<long long*>(bcost + T);                          // [T]
  double* rows = reinterpret_cast<double*>(bidx + T);                                 // [kPoolRows][T]
  long long* idx_row = reinterpret_cast<long long*>(rows + kPoolRows * T);            // [T]
  double* pool_pl = reinterpret_cast<double*>(idx_row + T);                           // [kWarps][6][kPoolCap]: planes
  int* pool_ok = reinterpret_cast<int*>(pool_pl + kWarps * 6 * kPoolCap);             // [kWarps][kPoolCap]: owner << 16 | k
  int* queue = pool_ok + kWarps * kPoolCap;                                            // [kWarps][kQueueCap]: owner << 16 | k
  unsigned* res = reinterpret_cast<unsigned*>(queue + kWarps * kQueueCap);             // [T]: min (k << 2 | verdict)
  static_assert((kWarps * 8 * 8 + T * 16 + kPoolRows * T * 8 + T * 8 + kWarps * kPoolCap * (6 * 8 + 4) + kWarps * kQueueCap * 4 + T * 4) %
                        128 == 0,
                "the table (TMA destination) starts on a 128-byte boundary");
  double* tab = reinterpret_cast<double*>(res + T);                                   // [mp][stride]
  int* oflags = reinterpret_cast<int*>(tab + (size_t)p.stride * mp);                  // [mp]
  unsigned* inbits = reinterpret_cast<unsigned*>(oflags + mp);                        // [(mp + 31) / 32][T], static tables
  __shared__ __align__(8) uint64_t mbar;

  const uint32_t blob_bytes = (uint32_t)(((size_t)p.stride * mp) * 8 + (size_t)mp * 4);
  if (tid == 0) mbar_init(&mbar, 1);
  if (tid < kWarps * 8) wcnt[tid] = 0ull;
  bcost[tid] = __longlong_as_double(0x7ff0000000000000ll);
  bidx[tid] = -1;
#pragma unroll
  for (int j = 0; j < kPoolRows; j++) rows[j * T + tid] = (j == PR_PL + 3) ? 1.0 : 0.0;  // a harmless plane for idle lanes
  __syncthreads();
  if (blob_bytes > 0) {
    if (tid == 0) tma_bulk_g2s(tab, p.obs_blob, blob_bytes, &mbar);
    mbar_wait(&mbar, 0);
  }

  double c[18];
#pragma unroll
  for (int j = 0; j < 18; j++) c[j] = 0.0;
  double* my_row = rows + tid;
  double* wpool_pl = pool_pl + warp * 6 * kPoolCap;
  int* wpool_ok = pool_ok + warp * kPoolCap;
  int* wqueue = queue + warp * kQueueCap;
  const unsigned lt_mask = (1u << lane) - 1u;
  double2 stack[kMaxDepth];  // the running task's parked low children {ts, lo_tf}
  unsigned long long avail = 0;
  bool exhausted = false;
  const bool early = PLAIN ? false : (p.early_exit != 0);
  uint8_t* const vmat = PLAIN ? nullptr : p.verdict_matrix;

  // with an accept list (input filter) the work items are positions of the list, else the candidates themselves
  const long long n_work = p.n_list ? (long long)*p.n_list : p.n;
  while (!exhausted) {
    // ---- a batch of 32 candidates
    unsigned long long b = 0;
    if (lane == 0) b = atomicAdd(p.work_counter, 32ull);
    b = __shfl_sync(kFull, b, 0);
    if (b >= (unsigned long long)n_work) break;
    if (b + 32ull >= (unsigned long long)n_work) exhausted = true;
    if (p.progress) {  // streaming input: wait until the copy engine has delivered the whole batch
      unsigned long long want = b + 32ull;
      if (want > (unsigned long long)p.n) want = (unsigned long long)p.n;
      if (!wait_for_input(p, want, avail, lane)) break;
    }
    const long long pos = (long long)b + lane;
    const long long idx = (p.list && pos < n_work) ? p.list[pos] : pos;
    const bool have = pos < n_work && !(p.skip && p.skip[idx] != 0);  // planner: dropped by an earlier stage
    double t0 = 0.0, t1 = 0.0, cost = 0.0;
    V3 x0 = v3(0, 0, 0), x1 = v3(0, 0, 0), xm0 = v3(0, 0, 0);
    double hull_r = __longlong_as_double(0x7ff0000000000000ll);  // bound of |X(t) - X(mid)| on the window; +inf: no cull
    if (have) {
      load_trajectory_compact<MODE>(p, idx, c, t0, t1, cost);
      if (MODE == MODE_BC) {
        if (p.cost) p.cost[idx] = cost;
        if (p.coeffs) {
#pragma unroll
          for (int j = 0; j < 18; j++) p.coeffs[j * p.coeff_stride + idx] = c[j];
        }
      }
      if (!MOVING) {
        x0 = traj_value(c, t0), x1 = traj_value(c, t1);
        const V3 xm = traj_value(c, (t0 + t1) / 2);  // the first section's midpoint (CollisionChecker.cpp:87), every obstacle
        my_row[(PR_XM + 0) * T] = xm.x, my_row[(PR_XM + 1) * T] = xm.y, my_row[(PR_XM + 2) * T] = xm.z;
        const double amax0 = side_filter_bound(c, t0, t1);
        my_row[PR_AMAX * T] = amax0;
        xm0 = xm;
        // a window shorter than minT makes every first section Indeterminable (:93-96): no cull then
        if (!(t1 - t0 < p.min_t)) hull_r = traj_hull_radius(c, t0, t1, xm, amax0);
      }
      my_row[PR_T0 * T] = t0;
      my_row[PR_T1 * T] = t1;
#pragma unroll
      for (int j = 0; j < 18; j++) my_row[(PR_C + j) * T] = c[j];
    }
    idx_row[tid] = idx;
    res[tid] = 0xffffffffu;
    if (!MOVING) {
      // CollisionCheck's end-point tests (CollisionChecker.cpp:73-77) against the whole table: lane = trajectory, the
      // obstacle record is a broadcast load; bit k of the lane's words = an end point is inside obstacle k
      for (int k0 = 0; k0 < m; k0 += 32) {
        unsigned bits = 0;
        const int kend = min(32, m - k0);
        for (int j = 0; j < kend; j++) {
          TableObs o;
          o.rec = tab + (size_t)(k0 + j) * p.stride;
          bits |= ends_inside(o, x0, x1) ? (1u << j) : 0u;
        }
        inbits[(k0 >> 5) * T + tid] = have ? bits : 0u;
      }
    }
    __syncwarp();

    // ---- the batch. Lane state of a running task: owner, obstacle, interval, parked low children.
    int st = LS_IDLE, owner = lane, kk = 0, sp = 0, pool_n = 0, k = 0, q_n = 0;
    double ts = 0.0, tf = 0.0;
    unsigned n1 = 0, n2 = 0, n3 = 0;  // pairs this lane decided as Collision / Indeterminable / depth cap (all-pairs mode)
    for (;;) {
      // early exit: a task of an obstacle behind the owner's known first hit is moot
      if (early && st != LS_IDLE && (res[wbase + owner] >> 2) < (unsigned)kk) st = LS_IDLE;
      // ---- idle lanes take tasks from the top of the pool; a pooled task has its plane and waits for the solver
      {
        const unsigned idle_m = __ballot_sync(kFull, st == LS_IDLE);
        const int take_n = min(__popc(idle_m), pool_n);
        if (take_n > 0) {
          const int rank = __popc(idle_m & lt_mask);
          if (st == LS_IDLE && rank < take_n) {
            const int slot = pool_n - 1 - rank;
            const int ok = wpool_ok[slot];
            owner = ok >> 16;
            kk = ok & 0xffff;
#pragma unroll
            for (int j = 0; j < 6; j++) my_row[(PR_PL + j) * T] = wpool_pl[j * kPoolCap + slot];
            const double* r_ = rows + (wbase + owner);
            ts = r_[PR_T0 * T], tf = r_[PR_T1 * T];
            if (MOVING && (oflags[kk] & OBS_MOVING)) {
              const double* rec = tab + (size_t)kk * p.stride;
              const double ms = rec[OF_MT0], me = rec[OF_MT1];
              ts = (ts < ms) ? ms : ts;  // std::max(_startTime, rhs start), Trajectory.hpp:121-128
              tf = (me < tf) ? me : tf;  // std::min(_endTime, rhs end)
            }
            sp = 0;
            st = LS_NEED_SOLVE;
          }
          pool_n -= take_n;
          __syncwarp();
        }
      }
      const int n_plane = __popc(__ballot_sync(kFull, st == LS_NEED_PLANE));
      const int n_solve = __popc(__ballot_sync(kFull, st == LS_NEED_SOLVE));
      // ---- the cull pass feeds the queue of first sections: lane = own trajectory, obstacle k = broadcast loads. A pair
      // whose end-point bit is set is a Collision without a section (:73-77); a pair whose obstacle is far from the whole
      // trajectory (far_from_obstacle) is NoCollision without a plane; the others wait for an A trip.
      while (k < m && q_n <= kQueueCap - 32) {
        const double* rec = tab + (size_t)k * p.stride;
        bool push = have && !(early && (res[tid] >> 2) < (unsigned)k);
        if (!MOVING) {
          const bool bit = (inbits[(k >> 5) * T + tid] >> (k & 31)) & 1u;
          const bool far = far_from_obstacle(xm0, hull_r, v3(rec[OF_CX], rec[OF_CY], rec[OF_CZ]), rec[OF_RBOUND]);
          if (push & bit) {
            if (vmat) vmat[idx * (long long)m + k] = (uint8_t)1;
            n1++;
            atomicMin(&res[tid], ((unsigned)k << 2) | 1u);
          } else if (push & far) {
            if (vmat) vmat[idx * (long long)m + k] = (uint8_t)0;
          }
          push = push & !bit & !far;
        }
        const unsigned push_m = __ballot_sync(kFull, push);
        if (push) wqueue[q_n + __popc(push_m & lt_mask)] = (lane << 16) | k;
        q_n += __popc(push_m);
        k++;
      }
      __syncwarp();
      const bool a_ok = q_n > 0 && (q_n >= 32 || k >= m) && pool_n <= kPoolCap - 32;
      int trip;  // 0: A, 1: P, 2: S
      if (n_solve >= p.th_solve) trip = 2;
      else if (n_plane >= p.th_plane) trip = 1;
      else if (a_ok) trip = 0;
      else if (n_solve > 0 && n_solve >= n_plane) trip = 2;
      else if (n_plane > 0) trip = 1;
      else if (q_n > 0) trip = 0;
      else break;  // table through, queue and pool empty, no lane holds a task: the batch is done

      int pv = -1;  // verdict of the task that ended in this trip, if any
      // ---- every trip works on ONE section per lane: (trajectory of lane `o_lane`, obstacle `o_k`, [f_ts, f_tf]).
      // A trip: the first section of (own trajectory, obstacle k); P and S trips: the lane's task.
      const bool trip_a = trip == 0;
      int a_lane = lane, a_k = 0;
      bool a_have = false;
      if (trip_a) {  // every lane takes a first section from the top of the queue
        const int take_n = min(32, q_n);
        a_have = lane < take_n;
        const int e = wqueue[a_have ? q_n - 1 - lane : 0];
        a_lane = e >> 16, a_k = e & 0xffff;
        q_n -= take_n;
      }
      const int o_lane = trip_a ? a_lane : owner, o_k = trip_a ? a_k : kk;
      const double* orow = rows + (wbase + o_lane);
      const double* rec = tab + (size_t)o_k * p.stride;
#pragma unroll
      for (int j = 0; j < 18; j++) c[j] = orow[(PR_C + j) * T];
      const double o_t0 = orow[PR_T0 * T], o_t1 = orow[PR_T1 * T];
      double f_ts = trip_a ? o_t0 : ts, f_tf = trip_a ? o_t1 : tf;
      bool active = trip_a ? (a_have && !(early && (res[wbase + a_lane] >> 2) < (unsigned)a_k))
                           : (st == (trip == 1 ? LS_NEED_PLANE : LS_NEED_SOLVE));
      int pre = -1;  // A trip, moving obstacle: decided before the section: 1 an end point is inside, 0 no common window
      if (MOVING) {
        // Trajectory::operator- (Trajectory.hpp:121-128): coefficient-wise difference (zeros for a static obstacle of the
        // table) on the common window
#pragma unroll
        for (int j = 0; j < 18; j++) c[j] = c[j] - rec[OF_MOTION + j];
        if (trip_a) {
          if (oflags[o_k] & OBS_MOVING) {
            const double ms = rec[OF_MT0], me = rec[OF_MT1];
            f_ts = (o_t0 < ms) ? ms : o_t0;  // std::max(_startTime, rhs start)
            f_tf = (me < o_t1) ? me : o_t1;  // std::min(_endTime, rhs end)
            if (active & !(f_ts <= f_tf)) pre = 0;  // no common time (the reference would assert, Trajectory.hpp:68)
          }
          TableObs o;
          o.rec = rec;
          const bool in_ends = ends_inside(o, traj_value(c, f_ts), traj_value(c, f_tf));  // :73-77
          __syncwarp();
          if (active & (pre < 0) & in_ends) pre = 1;
        }
      }
      active = active & (pre < 0);
      const double amax = MOVING ? side_filter_bound(c, f_ts, f_tf) : orow[PR_AMAX * T];
      const double mid = (f_ts + f_tf) / 2;  // :87

      if (trip != 2) {
        // ---- the PLANE stage. X(mid) of a static first section was evaluated when the batch was built.
        V3 xm;
        if (!MOVING && trip_a) xm = v3(orow[(PR_XM + 0) * T], orow[(PR_XM + 1) * T], orow[(PR_XM + 2) * T]);
        else xm = traj_value(c, mid);
        TableObs o;
        o.rec = rec;
        V3 pp, pn;
        const int r = plane_stage(c, o, xm, f_ts, f_tf, amax, p.min_t, active, pp, pn);
        if (trip_a) {
          // a first section: decided, or a task for the pool
          const bool spawn = active & (r == PS_SOLVE);
          const unsigned spawn_m = __ballot_sync(kFull, spawn);
          if (spawn) {
            const int slot = pool_n + __popc(spawn_m & lt_mask);
            wpool_ok[slot] = (a_lane << 16) | a_k;
            wpool_pl[0 * kPoolCap + slot] = pp.x, wpool_pl[1 * kPoolCap + slot] = pp.y, wpool_pl[2 * kPoolCap + slot] = pp.z;
            wpool_pl[3 * kPoolCap + slot] = pn.x, wpool_pl[4 * kPoolCap + slot] = pn.y, wpool_pl[5 * kPoolCap + slot] = pn.z;
          }
          const int v = (pre >= 0) ? pre : (active ? (r & 3) : -1);  // PS_SOLVE & 3 == 0: NoCollision until the task says otherwise
          if (v >= 0) {
            if (vmat) vmat[idx_row[wbase + a_lane] * (long long)m + a_k] = (uint8_t)v;
            if (v != 0) {
              n1 += (v == 1), n2 += (v == 2);
              atomicMin(&res[wbase + a_lane], ((unsigned)a_k << 2) | (unsigned)v);
            }
          }
          pool_n += __popc(spawn_m);
          __syncwarp();
        } else if (active) {
          // a child section of the lane's task
          if (r == PS_SOLVE) {
            my_row[(PR_PL + 0) * T] = pp.x, my_row[(PR_PL + 1) * T] = pp.y, my_row[(PR_PL + 2) * T] = pp.z;
            my_row[(PR_PL + 3) * T] = pn.x, my_row[(PR_PL + 4) * T] = pn.y, my_row[(PR_PL + 5) * T] = pn.z;
            st = LS_NEED_SOLVE;
          } else if (r == PS_CLEAR && sp > 0) {  // subtree clear: resume the innermost parked low child
            const double2 e = stack[--sp];
            ts = e.x, tf = e.y;
          } else {
            pv = r;  // NoCollision with nothing parked, Collision or Indeterminable: the pair is done
          }
        }
      } else {
        // ---- the SOLVE stage
        const V3 pp = v3(my_row[(PR_PL + 0) * T], my_row[(PR_PL + 1) * T], my_row[(PR_PL + 2) * T]);
        const V3 pn = v3(my_row[(PR_PL + 3) * T], my_row[(PR_PL + 4) * T], my_row[(PR_PL + 5) * T]);
        double pc[5], r0, r1, r2, r3;
        int n;
        {
          FastMath f;
          n = critical_times(f, c, pn, pc, r0, r1, r2, r3);
          if (!warp_all(f.ok | !active)) {
            const PlainRoots q = critical_times_plain(orow, MOVING ? rec + OF_MOTION : nullptr, T, pn);
            n = q.n, r0 = q.r[0], r1 = q.r[1], r2 = q.r[2], r3 = q.r[3];
#pragma unroll
            for (int j = 0; j < 5; j++) pc[j] = q.pc[j];
          }
        }
        Children ch;
        section_children(c, ts, tf, mid, pp, pn, pc, n, r0, r1, r2, r3, amax, active, ch);
        if (active) {
          if (ch.has_hi | ch.has_lo) {
            if (ch.has_hi & ch.has_lo) {  // the parent's low tail call waits until the high subtree is clear
              if (sp < kMaxDepth) stack[sp++] = make_double2(ts, ch.lo_tf);
              else pv = 3;  // deeper than RQCD_MAX_DEPTH: never silent truncation
            }
            if (ch.has_hi) ts = ch.t_next;
            else tf = ch.t_next;
            st = LS_NEED_PLANE;
          } else if (sp > 0) {
            const double2 e = stack[--sp];
            ts = e.x, tf = e.y;
            st = LS_NEED_PLANE;
          } else {
            pv = 0;
          }
        }
      }
      if (pv >= 0) {
        st = LS_IDLE;
        if (pv != 0) {
          n1 += (pv == 1), n2 += (pv == 2), n3 += (pv == 3);
          atomicMin(&res[wbase + owner], ((unsigned)kk << 2) | (unsigned)pv);
          if (vmat) vmat[idx_row[wbase + owner] * (long long)m + kk] = (uint8_t)pv;
        }
      }
    }

    // ---- the batch is through: counters per warp, verdicts per owner
    __syncwarp();
    const unsigned word = res[tid];
    const bool hit = have && word != 0xffffffffu;
    const int verdict = hit ? (int)(word & 3u) : 0;
    const int first = hit ? (int)(word >> 2) : -1;
    unsigned c0;  // NoCollision pairs of this lane's trajectory
    if (early) {  // the driver's loop stops at the first hit: pairs behind it were not executed
      n1 = verdict == 1, n2 = verdict == 2, n3 = verdict == 3;
      c0 = have ? (hit ? (unsigned)first : (unsigned)m) : 0u;
    } else {
      c0 = have ? (unsigned)m : 0u;  // minus the s1 + s2 + s3 below
    }
    const unsigned s0 = __reduce_add_sync(kFull, c0);
    const unsigned s1 = __reduce_add_sync(kFull, n1), s2 = __reduce_add_sync(kFull, n2), s3 = __reduce_add_sync(kFull, n3);
    const int tv = have ? verdict : -1;
    const unsigned q0 = __ballot_sync(kFull, tv == 0), q1 = __ballot_sync(kFull, tv == 1);
    const unsigned q2 = __ballot_sync(kFull, tv == 2), q3 = __ballot_sync(kFull, tv == 3);
    if (lane == 0) {
      unsigned long long* w = wcnt + warp * 8;
      w[0] += __popc(q0), w[1] += __popc(q1), w[2] += __popc(q2), w[3] += __popc(q3);
      w[4] += early ? s0 : s0 - s1 - s2 - s3;
      w[5] += s1, w[6] += s2, w[7] += s3;
    }
    if (have) {
      if (p.verdict) p.verdict[idx] = (uint8_t)verdict;
      if (p.first_obstacle) p.first_obstacle[idx] = first;
      if (verdict == 0 && (cost < bcost[tid] || (cost == bcost[tid] && bidx[tid] >= 0 && p.index_base + idx < bidx[tid]))) {  // lowest index on ties
        bcost[tid] = cost;
        bidx[tid] = p.index_base + idx;
      }
    }
    __syncwarp();
  }

  // ---- CTA partial
  __syncthreads();
  if (tid < 32) {
    double bc = __longlong_as_double(0x7ff0000000000000ll);
    long long bi = -1;
    for (int j = lane; j < T; j += 32) {
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
    if (lane < 8) {
      unsigned long long sum = 0;
      for (int w = 0; w < kWarps; w++) sum += wcnt[w * 8 + lane];
      if (lane < 4) out->traj[lane] = sum;
      else out->pair[lane - 4] = sum;
    }
    if (lane == 0) {
      out->best_cost = bc;
      out->best_index = bi;
    }
  }
}

__device__ double g_pool_rcp_const[3];
__global__ void k_pool_init_constants() {
  g_pool_rcp_const[0] = rcp_refined(3.0);
  g_pool_rcp_const[1] = rcp_refined(9.0);
  g_pool_rcp_const[2] = rcp_refined(54.0);
}

template <int T, bool PLAIN>
const void* pool_fn(int mode, bool moving) {
  if (moving)
    return mode == MODE_BC ? reinterpret_cast<const void*>(&k_pool<MODE_BC, true, T, PLAIN>)
                           : reinterpret_cast<const void*>(&k_pool<MODE_COEFF, true, T, PLAIN>);
  return mode == MODE_BC ? reinterpret_cast<const void*>(&k_pool<MODE_BC, false, T, PLAIN>)
                         : reinterpret_cast<const void*>(&k_pool<MODE_COEFF, false, T, PLAIN>);
}
template <bool PLAIN>
const void* pool_fn(int mode, bool moving, int threads) {
  if (threads == 256) return pool_fn<256, PLAIN>(mode, moving);
  if (threads == 384) return pool_fn<384, PLAIN>(mode, moving);
  return pool_fn<512, PLAIN>(mode, moving);
}
const void* pool_fn(int mode, bool moving, int threads, bool plain) {
  return plain ? pool_fn<true>(mode, moving, threads) : pool_fn<false>(mode, moving, threads);
}

}  // namespace

// this translation unit's copy of the refined reciprocals of solveP3's constant divisors (see init_device_constants)
cudaError_t init_pool_constants(cudaStream_t s) {
  k_pool_init_constants<<<1, 1, 0, s>>>();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  void* src = nullptr;
  e = cudaGetSymbolAddress(&src, g_pool_rcp_const);
  if (e != cudaSuccess) return e;
  e = cudaMemcpyToSymbolAsync(kRcpConst, src, sizeof(double) * 3, 0, cudaMemcpyDeviceToDevice, s);
  if (e != cudaSuccess) return e;
  return cudaStreamSynchronize(s);
}

size_t pool_smem_bytes(int mp, int stride, int threads) {
  const size_t T = (size_t)threads, W = T / 32;
  size_t u = (size_t)stride * mp * 8 + (size_t)mp * 4;  // table + flags (mp % 4 == 0 keeps 16-byte alignment)
  u += W * 8 * 8 + T * 16;                              // warp counters, best cost / index
  u += T * 8 * kPoolRows + T * 8;                       // trajectory rows, global indices
  u += W * kPoolCap * (6 * 8 + 4) + W * kQueueCap * 4;  // task pools (planes, owner/obstacle words), first-section queues
  u += T * 4 + (size_t)((mp + 31) / 32) * T * 4;        // verdict words, end-point bits
  return u;
}

// CTA size for a table: two CTAs of 256 threads per SM when their shared memory fits twice, else one CTA of 512 or
// 384 threads (the table is then shared by all warps of the SM). 0: the table does not fit beside the rows.
int pool_threads(int mp, int stride, size_t smem_per_sm, size_t smem_per_block_max) {
  const size_t reserve = 1024;  // per resident CTA
  if (2 * (pool_smem_bytes(mp, stride, 256) + reserve) <= smem_per_sm) return 256;
  if (pool_smem_bytes(mp, stride, 512) <= smem_per_block_max) return 512;
  if (pool_smem_bytes(mp, stride, 384) <= smem_per_block_max) return 384;
  if (pool_smem_bytes(mp, stride, 256) <= smem_per_block_max) return 256;
  return 0;
}

const char* pool_kernel_name(bool moving) { return moving ? "k_pool<moving>" : "k_pool<static>"; }

cudaError_t pool_prepare(int mode, bool moving, int threads, size_t smem) {
  cudaError_t e = cudaFuncSetAttribute(pool_fn(mode, moving, threads, true), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(pool_fn(mode, moving, threads, false), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

int pool_max_blocks_per_sm(int mode, bool moving, int threads, size_t smem) {
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, pool_fn(mode, moving, threads, false), threads, smem) != cudaSuccess) return 0;
  return nb;
}

cudaError_t launch_pool(const CheckParams& p, int mode, bool moving, int threads, int grid, size_t smem, cudaStream_t s) {
  CheckParams q = p;
  void* args[] = {&q};
  const bool plain = p.verdict_matrix == nullptr && p.early_exit == 0;
  return cudaLaunchKernel(pool_fn(mode, moving, threads, plain), dim3(grid), dim3(threads), args, smem, s);
}

}  // namespace rqcd
