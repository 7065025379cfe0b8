/*
 * oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).  Plain C99 restatement of the reference
 * algorithm; every function cites the reference file:line it follows (paths relative to the
 * reference root).  Build: gcc -O2 -ffp-contract=off (no FMA contraction, like the reference's
 * x86-64 -O3 build without -march).  Nothing in the product path may call into this file.
 */
#define _GNU_SOURCE
#include "oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * Quartic::solveP3 -- src/quartic.cpp:39-72.  eps, M_2PI: include/Quartic/quartic.h:34-36.
 * ---------------------------------------------------------------------------------------------- */
static const double ORC_EPS = 1e-12;
static const double ORC_2PI = 2 * 3.141592653589793238463L;

typedef struct solve_info {
  int trig;  /* 1 if the three-real-root branch ran */
  int pairs; /* real root pairs appended by the quartic */
} solve_info;

static int solve_p3(double a, double b, double c, double* x, solve_info* si) {
  double a2 = a * a;
  double q = (a2 - 3 * b) / 9;
  double r = (a * (2 * a2 - 9 * b) + 27 * c) / 54;
  double r2 = r * r;
  double q3 = q * q * q;
  double A, B;
  if (r2 < q3) { /* :46-57 */
    double t = r / sqrt(q3);
    if (t < -1) t = -1;
    if (t > 1) t = 1;
    t = acos(t);
    a /= 3;
    q = -2 * sqrt(q);
    x[0] = q * cos(t / 3) - a;
    x[1] = q * cos((t + ORC_2PI) / 3) - a;
    x[2] = q * cos((t - ORC_2PI) / 3) - a;
    if (si) si->trig = 1;
    return 3;
  } else { /* :58-71 */
    A = -pow(fabs(r) + sqrt(r2 - q3), 1. / 3);
    if (r < 0) A = -A;
    B = (0 == A ? 0 : q / A);
    a /= 3;
    x[0] = (A + B) - a;
    x[1] = -0.5 * (A + B) - a;
    x[2] = 0.5 * sqrt(3.) * (A - B);
    if (si) si->trig = 0;
    if (fabs(x[2]) < ORC_EPS) {
      x[2] = x[1];
      return 2;
    }
    return 1;
  }
}

/* Quartic::solve_quartic -- src/quartic.cpp:77-156. */
static int solve_quartic(double a, double b, double c, double d, double* root, solve_info* si) {
  double a3 = -b; /* :81-83 resolvent y^3 - b y^2 + (ac-4d) y - a^2 d - c^2 + 4bd */
  double b3 = a * c - 4. * d;
  double c3 = -a * a * d - c * c + 4. * b * d;
  int rCnt = 0;
  double x3[3];
  int iZeroes = solve_p3(a3, b3, c3, x3, si); /* :91 */
  double q1, q2, p1, p2, D, sqD, y;
  y = x3[0];
  if (iZeroes != 1) { /* :97-101 choose y of maximal magnitude */
    if (fabs(x3[1]) > fabs(y)) y = x3[1];
    if (fabs(x3[2]) > fabs(y)) y = x3[2];
  }
  D = y * y - 4 * d;       /* :105 */
  if (fabs(D) < ORC_EPS) { /* :106-120 */
    q1 = q2 = y * 0.5;
    D = a * a - 4 * (b - y);
    if (fabs(D) < ORC_EPS) {
      p1 = p2 = a * 0.5;
    } else {
      sqD = sqrt(D);
      p1 = (a + sqD) * 0.5;
      p2 = (a - sqD) * 0.5;
    }
  } else { /* :121-129 */
    sqD = sqrt(D);
    q1 = (y + sqD) * 0.5;
    q2 = (y - sqD) * 0.5;
    p1 = (a * q1 - c) / (q1 - q2);
    p2 = (c - a * q2) / (q1 - q2);
  }
  D = p1 * p1 - 4 * q1; /* :133-142 */
  if (!(D < 0.0)) {
    sqD = sqrt(D);
    root[rCnt] = (-p1 + sqD) * 0.5;
    ++rCnt;
    root[rCnt] = (-p1 - sqD) * 0.5;
    ++rCnt;
  }
  D = p2 * p2 - 4 * q2; /* :145-153 */
  if (!(D < 0.0)) {
    sqD = sqrt(D);
    root[rCnt] = (-p2 + sqD) * 0.5;
    ++rCnt;
    root[rCnt] = (-p2 - sqD) * 0.5;
    ++rCnt;
  }
  if (si) si->pairs = rCnt / 2;
  return rCnt;
}

int orc_solve_p3(double a, double b, double c, double* x) { return solve_p3(a, b, c, x, NULL); }
int orc_solve_quartic(double a, double b, double c, double d, double* root) {
  return solve_quartic(a, b, c, d, root, NULL);
}

void orc_solve_cubic_batch(const double* coef, int64_t n, double* roots, int32_t* count) {
  for (int64_t i = 0; i < n; i++) {
    double x[3];
    count[i] = solve_p3(coef[i], coef[n + i], coef[2 * n + i], x, NULL);
    for (int k = 0; k < 3; k++) roots[k * n + i] = x[k];
  }
}

void orc_solve_quartic_batch(const double* coef, int64_t n, double* roots, int32_t* count) {
  for (int64_t i = 0; i < n; i++) {
    double x[4] = {NAN, NAN, NAN, NAN};
    count[i] = solve_quartic(coef[i], coef[n + i], coef[2 * n + i], coef[3 * n + i], x, NULL);
    for (int k = 0; k < 4; k++) roots[k * n + i] = x[k];
  }
}

/* ------------------------------------------------------------------------------------------------
 * Vec3 / Trajectory / obstacles
 * ---------------------------------------------------------------------------------------------- */
typedef struct v3 {
  double x, y, z;
} v3;

static inline v3 v3_sub(v3 a, v3 b) { return (v3){a.x - b.x, a.y - b.y, a.z - b.z}; } /* Vec3.hpp:125-127 */
static inline v3 v3_add(v3 a, v3 b) { return (v3){a.x + b.x, a.y + b.y, a.z + b.z}; } /* Vec3.hpp:122-124 */
static inline double v3_dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; } /* Vec3.hpp:96-98 */
static inline double v3_norm(v3 a) { return sqrt(v3_dot(a, a)); }                      /* Vec3.hpp:107-114 */
/* Vec3::GetUnitVector narrows the norm to float before dividing -- Vec3.hpp:117-120. */
static inline v3 v3_unit(v3 a) {
  float const n = (float)v3_norm(a);
  double nd = n;
  return (v3){a.x / nd, a.y / nd, a.z / nd};
}

/* Trajectory::GetValue, not Horner, left to right per component -- Trajectory.hpp:76-82
 * with Vec3*double = three multiplies (Vec3.hpp:160-162). */
static inline v3 traj_value(const double* c, double t) {
  v3 r;
  r.x = c[0] * t * t * t * t * t + c[3] * t * t * t * t + c[6] * t * t * t + c[9] * t * t + c[12] * t + c[15];
  r.y = c[1] * t * t * t * t * t + c[4] * t * t * t * t + c[7] * t * t * t + c[10] * t * t + c[13] * t + c[16];
  r.z = c[2] * t * t * t * t * t + c[5] * t * t * t * t + c[8] * t * t * t + c[11] * t * t + c[14] * t + c[17];
  return r;
}

void orc_traj_value(const double coef[18], double t, double out[3]) {
  v3 r = traj_value(coef, t);
  out[0] = r.x;
  out[1] = r.y;
  out[2] = r.z;
}

/* Rotation::GetRotationMatrix -- Rotation.hpp:203-220. */
void orc_rotation_matrix(const double v[4], double R[9]) {
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

/* Rotation::Rotate -- Rotation.hpp:224-233. */
static inline v3 mat_mul(const double* R, v3 in) {
  return (v3){R[0] * in.x + R[1] * in.y + R[2] * in.z, R[3] * in.x + R[4] * in.y + R[5] * in.z,
              R[6] * in.x + R[7] * in.y + R[8] * in.z};
}

/* Obstacle with the per-obstacle constants hoisted (same expressions the reference re-evaluates on
 * every call: RectPrism.hpp:50 Inverse(), Rotation.hpp:63-65, :203-220). */
typedef struct prep_obs {
  int type;
  v3 c;
  double r;
  double half[3];
  double R[9], Ri[9];
} prep_obs;

static void prepare(const rqcd_obstacle* o, prep_obs* p) {
  p->type = o->type;
  p->c = (v3){o->center[0], o->center[1], o->center[2]};
  p->r = o->radius;
  for (int i = 0; i < 3; i++) p->half[i] = o->side[i] / 2; /* RectPrism.hpp:72-75 */
  if (o->type == RQCD_RECT_PRISM) {
    double qi[4] = {o->quat[0], -o->quat[1], -o->quat[2], -o->quat[3]}; /* Rotation.hpp:63-65 */
    orc_rotation_matrix(o->quat, p->R);
    orc_rotation_matrix(qi, p->Ri);
  } else {
    memset(p->R, 0, sizeof p->R);
    memset(p->Ri, 0, sizeof p->Ri);
  }
}

/* Sphere::IsPointInside Sphere.hpp:70-72; RectPrism::IsPointInside RectPrism.hpp:92-100. */
static inline int is_inside(const prep_obs* o, v3 tp) {
  if (o->type == RQCD_SPHERE) {
    return v3_norm(v3_sub(tp, o->c)) <= o->r;
  } else {
    v3 p = mat_mul(o->Ri, v3_sub(tp, o->c));
    return (fabs(p.x) <= o->half[0]) && (fabs(p.y) <= o->half[1]) && (fabs(p.z) <= o->half[2]);
  }
}

/* Sphere::GetTangentPlane Sphere.hpp:57-63; RectPrism::GetTangentPlane RectPrism.hpp:65-85. */
static inline void tangent_plane(const prep_obs* o, v3 tp, v3* point, v3* normal) {
  if (o->type == RQCD_SPHERE) {
    v3 u = v3_unit(v3_sub(tp, o->c));
    *normal = u;
    *point = v3_add((v3){u.x * o->r, u.y * o->r, u.z * o->r}, o->c);
  } else {
    v3 p = mat_mul(o->Ri, v3_sub(tp, o->c));
    double pl[3] = {p.x, p.y, p.z}, b[3];
    for (int i = 0; i < 3; i++) {
      if (pl[i] < -o->half[i])
        b[i] = -o->half[i];
      else if (pl[i] > o->half[i])
        b[i] = o->half[i];
      else
        b[i] = pl[i];
    }
    *point = v3_add(mat_mul(o->R, (v3){b[0], b[1], b[2]}), o->c);
    *normal = v3_unit(v3_sub(tp, *point));
  }
}

/* libstdc++ std::sort for n <= 16 is __insertion_sort (bits/stl_algo.h): replicated literally so that
 * NaN roots (src/quartic.cpp:110-118 can emit them) land where the reference puts them. */
static void sort_small(double* a, int n) {
  for (int i = 1; i < n; i++) {
    double val = a[i];
    if (val < a[0]) {
      for (int k = i; k > 0; k--) a[k] = a[k - 1];
      a[0] = val;
    } else {
      int j = i;
      while (val < a[j - 1]) {
        a[j] = a[j - 1];
        j--;
      }
      a[j] = val;
    }
  }
}

/* Statistics only (orc_stats.skip_*): are all Bernstein coefficients of D(t) = (X(t) - pp).pn on [ts, tf] positive?
 * D in powers of t, Taylor shift to ts, scale by h = tf - ts, power basis -> Bernstein basis of degree 5. */
static int bernstein_positive(const double* coef, double ts, double tf, v3 pp, v3 pn) {
  static const double binom5[6] = {1, 5, 10, 10, 5, 1};
  static const double binom[6][6] = {{1, 0, 0, 0, 0, 0}, {1, 1, 0, 0, 0, 0}, {1, 2, 1, 0, 0, 0},
                                     {1, 3, 3, 1, 0, 0}, {1, 4, 6, 4, 1, 0}, {1, 5, 10, 10, 5, 1}};
  double a[6], b[6];
  for (int j = 0; j < 6; j++) { /* a[j]: coefficient of t^j */
    const double* c = coef + 3 * (5 - j);
    a[j] = pn.x * c[0] + pn.y * c[1] + pn.z * c[2];
  }
  a[0] -= v3_dot(pp, pn);
  const double h = tf - ts;
  double hp = 1.0;
  for (int j = 0; j < 6; j++) {
    double s = 0.0, tp = 1.0;
    for (int k = j; k < 6; k++) {
      s += binom[k][j] * a[k] * tp;
      tp *= ts;
    }
    b[j] = s * hp;
    hp *= h;
  }
  for (int i = 0; i < 6; i++) {
    double beta = 0.0;
    for (int j = 0; j <= i; j++) beta += binom[i][j] / binom5[j] * b[j];
    if (!(beta > 1e-9)) return 0;
  }
  return 1;
}

/* CollisionChecker::CollisionCheckSection -- src/CollisionChecker.cpp:82-193. */
static int check_section(const double* coef, double ts, double tf, const prep_obs* obs, double minT,
                         orc_stats* st, int depth) {
  if (st) {
    st->sections++;
    if ((uint64_t)depth > st->max_depth) st->max_depth = depth;
  }
  double midTime = (ts + tf) / 2; /* :87 */
  v3 midpoint = traj_value(coef, midTime);
  if (is_inside(obs, midpoint)) return RQCD_COLLISION;  /* :89-92 */
  if (tf - ts < minT) return RQCD_INDETERMINABLE;       /* :93-96 */
  v3 pp, pn;
  tangent_plane(obs, midpoint, &pp, &pn); /* :99 */
  if (st) {
    st->tangent_sections++;
    if (depth > 1) st->child_tangent_sections++;
    if (bernstein_positive(coef, ts, tf, pp, pn)) {
      if (depth > 1) st->skip_child++; else st->skip_first++;
    }
  }

  /* :103-111 with Trajectory::GetDerivativeCoeffs (Trajectory.hpp:105-112) */
  double c[5] = {0, 0, 0, 0, 0};
  double nrm[3] = {pn.x, pn.y, pn.z};
  for (int dim = 0; dim < 3; dim++) {
    for (int k = 0; k < 5; k++) c[k] += nrm[dim] * ((double)(5 - k) * coef[3 * k + dim]);
  }
  double roots[4];
  int rootCount;
  solve_info si = {0, 0};
  if (fabs(c[0]) > (double)1e-6) { /* :116-121 */
    rootCount = solve_quartic(c[1] / c[0], c[2] / c[0], c[3] / c[0], c[4] / c[0], roots, &si);
  } else {
    rootCount = solve_p3(c[2] / c[1], c[3] / c[1], c[4] / c[1], roots, &si);
    if (st) st->cubic_fallback++;
  }
  if (st) {
    if (si.trig) st->trig++; else st->cardano++;
    st->root_pairs += si.pairs;
  }
  sort_small(roots, rootCount); /* :123 */
#ifdef ORC_TRACE /* debugging builds only (tools/oracle_perturbed) */
  if (orc_trace_on) {
    fprintf(stderr, "depth %d [%.17g, %.17g] mid %.17g n %d roots", depth, ts, tf, midTime, rootCount);
    for (int i = 0; i < rootCount; i++)
      fprintf(stderr, " %.17g (D %.3g)", roots[i], v3_dot(v3_sub(traj_value(coef, roots[i]), pp), pn));
    fprintf(stderr, " | D(ts) %.3g D(tf) %.3g trig %d\n", v3_dot(v3_sub(traj_value(coef, ts), pp), pn),
            v3_dot(v3_sub(traj_value(coef, tf), pp), pn), si.trig);
  }
#endif

  double lo[6], hi[6];
  int nlo = 0, nhi = 0;
  lo[nlo++] = ts;      /* :131 */
  hi[nhi++] = midTime; /* :132 */
  for (int i = 0; i < rootCount; i++) { /* :133-147 */
    if (roots[i] <= ts) {
      continue;
    } else if (roots[i] < midTime) {
      lo[nlo++] = roots[i];
    } else if (roots[i] < tf) {
      hi[nhi++] = roots[i];
    } else {
      break;
    }
  }
  lo[nlo++] = midTime; /* :148 */
  hi[nhi++] = tf;      /* :149 */

  for (int i = 1; i < nhi; i++) { /* :154-174 */
    if (st) st->test_points++;
    if (v3_dot(v3_sub(traj_value(coef, hi[i]), pp), pn) <= 0) {
      int res = check_section(coef, hi[i - 1], tf, obs, minT, st, depth + 1);
      if (res == RQCD_NO_COLLISION) break;
      return res;
    }
  }
  for (int i = nlo - 2; i >= 0; i--) { /* :177-191, reverse iteration */
    if (st) st->test_points++;
    if (v3_dot(v3_sub(traj_value(coef, lo[i]), pp), pn) <= 0) {
      return check_section(coef, ts, lo[i + 1], obs, minT, st, depth + 1);
    }
  }
  return RQCD_NO_COLLISION; /* :192 */
}

/* CollisionChecker::CollisionCheck(obstacle, minT) -- src/CollisionChecker.cpp:70-80. */
static int check_static(const double* coef, double t0, double t1, const prep_obs* obs, double minT,
                        orc_stats* st) {
  if (st) {
    st->checks++;
    if (obs->type == RQCD_SPHERE) st->sphere_checks++; else st->prism_checks++;
  }
  if (is_inside(obs, traj_value(coef, t0)) || is_inside(obs, traj_value(coef, t1))) return RQCD_COLLISION;
  return check_section(coef, t0, t1, obs, minT, st, 1);
}

/* Pair check incl. the moving-obstacle composition: Trajectory::operator- (Trajectory.hpp:121-128)
 * then CollisionCheck on the relative quintic.  Non-overlapping windows make the reference assert
 * (Trajectory.hpp:68); here they are NoCollision (no common time), stated in DESIGN.md. */
static int check_pair(const double* coef, double t0, double t1, const rqcd_obstacle* o, const prep_obs* p,
                      double minT, orc_stats* st) {
  if (!o->moving) return check_static(coef, t0, t1, p, minT, st);
  double rel[18];
  for (int k = 0; k < 18; k++) rel[k] = coef[k] - o->motion[k];
  double s = (t0 < o->motion_t_start) ? o->motion_t_start : t0; /* std::max(a,b) = a<b ? b : a */
  double e = (o->motion_t_end < t1) ? o->motion_t_end : t1;     /* std::min(a,b) = b<a ? b : a */
  if (!(s <= e)) {
    if (st) st->checks++;
    return RQCD_NO_COLLISION;
  }
  return check_static(rel, s, e, p, minT, st);
}

int orc_check_one(const double coef[18], double t_start, double t_end, const rqcd_obstacle* obs,
                  double min_time_section, orc_stats* stats) {
  prep_obs p;
  prepare(obs, &p);
  return check_pair(coef, t_start, t_end, obs, &p, min_time_section, stats);
}

/* ------------------------------------------------------------------------------------------------
 * CollisionChecker::CollisionCheck(Boundary) -- src/CollisionChecker.cpp:25-68, DOCUMENTED INTENT.
 * PARITY UNPINNED against the reference binary: the reference hands `roots` (not roots+2) to the solver, so
 * the prepended start/end times are overwritten and two uninitialised slots are read (:43-55); NaN roots
 * reach GetValue and trip its assert. What is restated here is what its comments describe: test the start
 * time, the end time and every root inside the window.
 * ---------------------------------------------------------------------------------------------- */
int orc_check_plane_one(const double coef[18], double t0, double t1, const double plane[6]) {
  v3 pp = {plane[0], plane[1], plane[2]};
  v3 pn = v3_unit((v3){plane[3], plane[4], plane[5]}); /* :29 */
  double c[5] = {0, 0, 0, 0, 0};
  double nrm[3] = {pn.x, pn.y, pn.z};
  for (int dim = 0; dim < 3; dim++) {
    for (int k = 0; k < 5; k++) c[k] += nrm[dim] * ((double)(5 - k) * coef[3 * k + dim]); /* :35-43 */
  }
  double roots[6];
  int rootCount;
  roots[0] = t0;
  roots[1] = t1;
  if (fabs(c[0]) > (double)1e-6) { /* :47-52, writing after the two end times */
    rootCount = solve_quartic(c[1] / c[0], c[2] / c[0], c[3] / c[0], c[4] / c[0], roots + 2, NULL);
  } else {
    rootCount = solve_p3(c[2] / c[1], c[3] / c[1], c[4] / c[1], roots + 2, NULL);
  }
  for (int i = 0; i < rootCount + 2; i++) { /* :54-66 */
    if (!(roots[i] >= t0 && roots[i] <= t1)) continue; /* outside the domain, or NaN */
    if (v3_dot(v3_sub(traj_value(coef, roots[i]), pp), pn) <= 0) return RQCD_COLLISION;
  }
  return RQCD_NO_COLLISION;
}

void orc_check_planes(const rqcd_coeff_soa* in, int64_t n, const double* planes, int32_t p, uint8_t* verdict,
                      uint8_t* matrix) {
  for (int64_t i = 0; i < n; i++) {
    double coef[18];
    for (int k = 0; k < 18; k++) coef[k] = in->coeffs[k * in->stride + i];
    double t0 = in->t_start ? in->t_start[i] : 0.0, t1 = in->t_end[i];
    int any = 0;
    for (int32_t k = 0; k < p; k++) {
      int v = orc_check_plane_one(coef, t0, t1, planes + 6 * k);
      if (matrix) matrix[i * p + k] = (uint8_t)v;
      any |= v;
    }
    if (verdict) verdict[i] = (uint8_t)any;
  }
}

/* ------------------------------------------------------------------------------------------------
 * SingleAxisTrajectory::GenerateTrajectory -- src/SingleAxisTrajectory.cpp:57-120
 * ---------------------------------------------------------------------------------------------- */
typedef struct axis_t {
  double p0, v0, a0, a, b, g, cost;
} axis_t;

static void gen_axis(double p0, double v0, double a0, double pf, double vf, double af, int posG, int velG,
                     int accG, double Tf, axis_t* out) {
  double delta_a = af - a0; /* :60-62 */
  double delta_v = vf - v0 - a0 * Tf;
  double delta_p = pf - p0 - v0 * Tf - 0.5 * a0 * Tf * Tf;
  const double T2 = Tf * Tf; /* :65-68 */
  const double T3 = T2 * Tf;
  const double T4 = T3 * Tf;
  const double T5 = T4 * Tf;
  double a, b, g;
  if (posG && velG && accG) { /* :71-76 */
    a = (60 * T2 * delta_a - 360 * Tf * delta_v + 720 * 1 * delta_p) / T5;
    b = (-24 * T3 * delta_a + 168 * T2 * delta_v - 360 * Tf * delta_p) / T5;
    g = (3 * T4 * delta_a - 24 * T3 * delta_v + 60 * T2 * delta_p) / T5;
  } else if (posG && velG) { /* :77-82 */
    a = (-120 * Tf * delta_v + 320 * delta_p) / T5;
    b = (72 * T2 * delta_v - 200 * Tf * delta_p) / T5;
    g = (-12 * T3 * delta_v + 40 * T2 * delta_p) / T5;
  } else if (posG && accG) { /* :83-88 */
    a = (-15 * T2 * delta_a + 90 * delta_p) / (2 * T5);
    b = (15 * T3 * delta_a - 90 * Tf * delta_p) / (2 * T5);
    g = (-3 * T4 * delta_a + 30 * T2 * delta_p) / (2 * T5);
  } else if (velG && accG) { /* :89-94 */
    a = 0;
    b = (6 * Tf * delta_a - 12 * delta_v) / T3;
    g = (-2 * T2 * delta_a + 6 * Tf * delta_v) / T3;
  } else if (posG) { /* :95-100 */
    a = 20 * delta_p / T5;
    b = -20 * delta_p / T4;
    g = 10 * delta_p / T3;
  } else if (velG) { /* :101-106 */
    a = 0;
    b = -3 * delta_v / T3;
    g = 3 * delta_v / T2;
  } else if (accG) { /* :107-112 */
    a = 0;
    b = 0;
    g = delta_a / Tf;
  } else { /* :113-116 */
    a = b = g = 0;
  }
  out->p0 = p0;
  out->v0 = v0;
  out->a0 = a0;
  out->a = a;
  out->b = b;
  out->g = g;
  /* :119 */
  out->cost = g * g + b * g * Tf + b * b * T2 / 3.0 + a * g * T2 / 3.0 + a * b * T3 / 4.0 + a * a * T4 / 20.0;
}

/* SingleAxisTrajectory.hpp:54-63 evaluators. */
static inline double ax_jerk(const axis_t* s, double t) { return s->g + s->b * t + (1 / 2.0) * s->a * t * t; }
static inline double ax_acc(const axis_t* s, double t) {
  return s->a0 + s->g * t + (1 / 2.0) * s->b * t * t + (1 / 6.0) * s->a * t * t * t;
}
static inline double ax_vel(const axis_t* s, double t) {
  return s->v0 + s->a0 * t + (1 / 2.0) * s->g * t * t + (1 / 6.0) * s->b * t * t * t +
         (1 / 24.0) * s->a * t * t * t * t;
}
static inline double ax_pos(const axis_t* s, double t) {
  return s->p0 + s->v0 * t + (1 / 2.0) * s->a0 * t * t + (1 / 6.0) * s->g * t * t * t +
         (1 / 24.0) * s->b * t * t * t * t + (1 / 120.0) * s->a * t * t * t * t * t;
}

/* RapidTrajectoryGenerator::Generate (.cpp:67-72) + GetCost (.hpp:214-216) + GetTrajectory (.hpp:232-241)
 * for trajectory i of a boundary-condition matrix. */
static void gen_traj(const rqcd_bc_soa* in, int64_t i, axis_t ax[3], double coef[18], double* cost, double* Tf_out) {
  const double* d = in->data;
  const int64_t s = in->stride;
  double Tf = d[RQCD_BC_TF * s + i];
  for (int k = 0; k < 3; k++) {
    gen_axis(d[(RQCD_BC_P0 + k) * s + i], d[(RQCD_BC_V0 + k) * s + i], d[(RQCD_BC_A0 + k) * s + i],
             d[(RQCD_BC_PF + k) * s + i], d[(RQCD_BC_VF + k) * s + i], d[(RQCD_BC_AF + k) * s + i],
             (in->goal_mask >> (3 * k)) & 1, (in->goal_mask >> (3 * k + 1)) & 1,
             (in->goal_mask >> (3 * k + 2)) & 1, Tf, &ax[k]);
  }
  *cost = ax[0].cost + ax[1].cost + ax[2].cost;
  for (int k = 0; k < 3; k++) {
    coef[0 + k] = ax[k].a / 120;
    coef[3 + k] = ax[k].b / 24;
    coef[6 + k] = ax[k].g / 6;
    coef[9 + k] = ax_acc(&ax[k], 0) / 2;
    coef[12 + k] = ax_vel(&ax[k], 0);
    coef[15 + k] = ax_pos(&ax[k], 0);
  }
  *Tf_out = Tf;
}

void orc_generate(const rqcd_bc_soa* in, int64_t n, rqcd_out* out, double* abg) {
  for (int64_t i = 0; i < n; i++) {
    axis_t ax[3];
    double coef[18], cost, Tf;
    gen_traj(in, i, ax, coef, &cost, &Tf);
    if (out && out->cost) out->cost[i] = cost;
    if (out && out->coeffs)
      for (int k = 0; k < 18; k++) out->coeffs[k * out->coeff_stride + i] = coef[k];
    if (abg)
      for (int k = 0; k < 3; k++) {
        abg[(0 + k) * n + i] = ax[k].a;
        abg[(3 + k) * n + i] = ax[k].b;
        abg[(6 + k) * n + i] = ax[k].g;
      }
  }
}

/* ------------------------------------------------------------------------------------------------
 * Batched drivers (the trial-loop bodies PerformanceEval.cpp:140-182, ForestPerformanceEval.cpp:186-225)
 * ---------------------------------------------------------------------------------------------- */
typedef struct job {
  int mode; /* 0 generate_check, 1 check, 2 pairwise */
  const rqcd_bc_soa* bc;
  const rqcd_coeff_soa* cf;
  const double* cost_in;
  const double* spheres;
  int64_t sphere_stride;
  int64_t n;
  const rqcd_obstacle* obs;
  const prep_obs* prep;
  int32_t m;
  double minT;
  uint32_t flags;
  int64_t index_base;
  rqcd_out* out;
  int want_stats;
} job;

typedef struct shard {
  const job* j;
  int64_t lo, hi;
  rqcd_summary sum;
  orc_stats st;
} shard;

static void summary_init(rqcd_summary* s) {
  memset(s, 0, sizeof *s);
  s->best_cost = INFINITY;
  s->best_index = -1;
}

static void stats_add(orc_stats* a, const orc_stats* b) {
  a->checks += b->checks;
  a->sections += b->sections;
  a->tangent_sections += b->tangent_sections;
  a->test_points += b->test_points;
  a->trig += b->trig;
  a->cardano += b->cardano;
  a->root_pairs += b->root_pairs;
  a->cubic_fallback += b->cubic_fallback;
  if (b->max_depth > a->max_depth) a->max_depth = b->max_depth;
  a->sphere_checks += b->sphere_checks;
  a->prism_checks += b->prism_checks;
  a->skip_first += b->skip_first;
  a->skip_child += b->skip_child;
  a->child_tangent_sections += b->child_tangent_sections;
}

static void* run_shard(void* arg) {
  shard* sh = (shard*)arg;
  const job* j = sh->j;
  rqcd_out* out = j->out;
  orc_stats* st = j->want_stats ? &sh->st : NULL;
  summary_init(&sh->sum);
  memset(&sh->st, 0, sizeof sh->st);
  for (int64_t i = sh->lo; i < sh->hi; i++) {
    double coef[18], cost = NAN, t0 = 0, t1;
    if (j->mode == 1) {
      for (int k = 0; k < 18; k++) coef[k] = j->cf->coeffs[k * j->cf->stride + i];
      t0 = j->cf->t_start ? j->cf->t_start[i] : 0.0;
      t1 = j->cf->t_end[i];
      if (j->cost_in) cost = j->cost_in[i];
    } else {
      axis_t ax[3];
      gen_traj(j->bc, i, ax, coef, &cost, &t1);
      if (out->coeffs)
        for (int k = 0; k < 18; k++) out->coeffs[k * out->coeff_stride + i] = coef[k];
    }
    if (out->cost && j->mode != 1) out->cost[i] = cost;
    int verdict = RQCD_NO_COLLISION, first = -1;
    if (j->mode == 2) {
      rqcd_obstacle o;
      memset(&o, 0, sizeof o);
      o.type = RQCD_SPHERE;
      for (int k = 0; k < 3; k++) o.center[k] = j->spheres[k * j->sphere_stride + i];
      o.radius = j->spheres[3 * j->sphere_stride + i];
      prep_obs p;
      prepare(&o, &p);
      verdict = check_static(coef, t0, t1, &p, j->minT, st);
      if (verdict != RQCD_NO_COLLISION) first = 0;
      sh->sum.pair_counts[verdict]++;
      sh->sum.pairs_checked++;
    } else {
      for (int32_t k = 0; k < j->m; k++) {
        int v = check_pair(coef, t0, t1, &j->obs[k], &j->prep[k], j->minT, st);
        sh->sum.pair_counts[v]++;
        sh->sum.pairs_checked++;
        if (out->verdict_matrix) out->verdict_matrix[i * j->m + k] = (uint8_t)v;
        if (v != RQCD_NO_COLLISION && first < 0) {
          verdict = v;
          first = k;
          if (j->flags & RQCD_F_EARLY_EXIT) break; /* ForestPerformanceEval.cpp:219-224 */
        }
      }
    }
    if (out->verdict) out->verdict[i] = (uint8_t)verdict;
    if (out->first_obstacle) out->first_obstacle[i] = first;
    sh->sum.traj_counts[verdict]++;
    if (verdict == RQCD_NO_COLLISION && cost < sh->sum.best_cost) {
      sh->sum.best_cost = cost;
      sh->sum.best_index = j->index_base + i;
    }
  }
  return NULL;
}

static void merge(const rqcd_summary* s, int n, rqcd_summary* m) {
  summary_init(m);
  for (int k = 0; k < n; k++) {
    for (int v = 0; v < 4; v++) {
      m->traj_counts[v] += s[k].traj_counts[v];
      m->pair_counts[v] += s[k].pair_counts[v];
    }
    m->pairs_checked += s[k].pairs_checked;
    if (s[k].best_index >= 0 &&
        (s[k].best_cost < m->best_cost ||
         (s[k].best_cost == m->best_cost && (m->best_index < 0 || s[k].best_index < m->best_index)))) {
      m->best_cost = s[k].best_cost;
      m->best_index = s[k].best_index;
    }
  }
}

static void run_job(job* j, int nthreads, orc_stats* stats) {
  rqcd_out dummy;
  memset(&dummy, 0, sizeof dummy);
  if (!j->out) j->out = &dummy;
  prep_obs* prep = NULL;
  if (j->m > 0) {
    prep = (prep_obs*)malloc(sizeof(prep_obs) * (size_t)j->m);
    for (int32_t k = 0; k < j->m; k++) prepare(&j->obs[k], &prep[k]);
  }
  j->prep = prep;
  j->want_stats = stats != NULL;
  if (nthreads < 1) nthreads = 1;
  if ((int64_t)nthreads > j->n) nthreads = j->n > 0 ? (int)j->n : 1;
  shard* sh = (shard*)calloc((size_t)nthreads, sizeof(shard));
  pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
  for (int t = 0; t < nthreads; t++) {
    sh[t].j = j;
    sh[t].lo = j->n * t / nthreads;
    sh[t].hi = j->n * (t + 1) / nthreads;
  }
  if (nthreads == 1) {
    run_shard(&sh[0]);
  } else {
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, run_shard, &sh[t]);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  }
  if (j->out->summary) {
    rqcd_summary* tmp = (rqcd_summary*)malloc(sizeof(rqcd_summary) * (size_t)nthreads);
    for (int t = 0; t < nthreads; t++) tmp[t] = sh[t].sum;
    merge(tmp, nthreads, j->out->summary);
    free(tmp);
  }
  if (stats) {
    memset(stats, 0, sizeof *stats);
    for (int t = 0; t < nthreads; t++) stats_add(stats, &sh[t].st);
  }
  free(sh);
  free(th);
  free(prep);
}

void orc_generate_check(const rqcd_bc_soa* in, int64_t n, const rqcd_obstacle* obs, int32_t m,
                        double min_time_section, uint32_t flags, int64_t index_base, rqcd_out* out,
                        int nthreads, orc_stats* stats) {
  job j;
  memset(&j, 0, sizeof j);
  j.mode = 0;
  j.bc = in;
  j.n = n;
  j.obs = obs;
  j.m = m;
  j.minT = min_time_section;
  j.flags = flags;
  j.index_base = index_base;
  j.out = out;
  run_job(&j, nthreads, stats);
}

void orc_check(const rqcd_coeff_soa* in, const double* cost, int64_t n, const rqcd_obstacle* obs, int32_t m,
               double min_time_section, uint32_t flags, int64_t index_base, rqcd_out* out, int nthreads,
               orc_stats* stats) {
  job j;
  memset(&j, 0, sizeof j);
  j.mode = 1;
  j.cf = in;
  j.cost_in = cost;
  j.n = n;
  j.obs = obs;
  j.m = m;
  j.minT = min_time_section;
  j.flags = flags;
  j.index_base = index_base;
  j.out = out;
  run_job(&j, nthreads, stats);
}

void orc_generate_check_pairwise(const rqcd_bc_soa* in, const double* spheres, int64_t sphere_stride,
                                 int64_t n, double min_time_section, int64_t index_base, rqcd_out* out,
                                 int nthreads, orc_stats* stats) {
  job j;
  memset(&j, 0, sizeof j);
  j.mode = 2;
  j.bc = in;
  j.spheres = spheres;
  j.sphere_stride = sphere_stride;
  j.n = n;
  j.m = 0;
  j.minT = min_time_section;
  j.index_base = index_base;
  j.out = out;
  run_job(&j, nthreads, stats);
}

/* ------------------------------------------------------------------------------------------------
 * Input feasibility -- src/RapidTrajectoryGenerator.cpp:74-160, src/SingleAxisTrajectory.cpp:131-197
 * (needed here to reproduce PerformanceEval's acceptance filter).
 * ---------------------------------------------------------------------------------------------- */
typedef struct feas_ctx {
  axis_t ax[3];
  double peak[3][2]; /* _accPeakTimes.t, SingleAxisTrajectory.cpp:133-158 */
  int peak_init[3];  /* _accPeakTimes.initialised: cleared only by Reset()/SetInitialState(), NOT by
                        GenerateTrajectory(), so a reused generator keeps stale peak times */
  double grav[3];
} feas_ctx;

static void peak_times(const axis_t* s, double t[2]) {
  if (s->a) {
    double det = s->b * s->b - 2 * s->g * s->a;
    if (det < 0) {
      t[0] = 0;
      t[1] = 0;
    } else {
      t[0] = (-s->b + sqrt(det)) / s->a;
      t[1] = (-s->b - sqrt(det)) / s->a;
    }
  } else {
    if (s->b)
      t[0] = -s->g / s->b;
    else
      t[0] = 0;
    t[1] = 0;
  }
}

static inline double dmin(double a, double b) { return (b < a) ? b : a; } /* std::min */
static inline double dmax(double a, double b) { return (a < b) ? b : a; } /* std::max */

static void minmax_acc(const axis_t* s, double pk[2], int* init, double* aMin, double* aMax, double t1, double t2) {
  if (!*init) { /* :133-158 */
    peak_times(s, pk);
    *init = 1;
  }
  *aMin = dmin(ax_acc(s, t1), ax_acc(s, t2)); /* :160-161 */
  *aMax = dmax(ax_acc(s, t1), ax_acc(s, t2));
  for (int i = 0; i < 2; i++) { /* :164-171 */
    if (pk[i] <= t1) continue;
    if (pk[i] >= t2) continue;
    *aMin = dmin(*aMin, ax_acc(s, pk[i]));
    *aMax = dmax(*aMax, ax_acc(s, pk[i]));
  }
}

static inline double sq(double x) { return x * x; } /* pow(x,2): g++ -O3 expands to x*x */

static double max_jerk_sq(const axis_t* s, double t1, double t2) { /* :182-197 */
  double jMaxSqr = dmax(sq(ax_jerk(s, t1)), sq(ax_jerk(s, t2)));
  if (s->a) {
    double tMax = -s->b / s->a;
    if (tMax > t1 && tMax < t2) jMaxSqr = dmax(sq(ax_jerk(s, tMax)), jMaxSqr);
  }
  return jMaxSqr;
}

static double thrust(const feas_ctx* f, double t) { /* RapidTrajectoryGenerator.hpp:192-194 */
  v3 a = {ax_acc(&f->ax[0], t) - f->grav[0], ax_acc(&f->ax[1], t) - f->grav[1], ax_acc(&f->ax[2], t) - f->grav[2]};
  return v3_norm(a);
}

/* work counters of the feasibility recursion (flop model of bench.py's `feas` block; single-threaded use only) */
static uint64_t g_feas_sections, g_feas_axis_iters, g_feas_full;
void orc_feas_counters(uint64_t out[3], int reset) {
  out[0] = g_feas_sections, out[1] = g_feas_axis_iters, out[2] = g_feas_full;
  if (reset) g_feas_sections = g_feas_axis_iters = g_feas_full = 0;
}

static int feas_section(feas_ctx* f, double fminA, double fmaxA, double wmaxA, double t1, double t2,
                        double minT) { /* RapidTrajectoryGenerator.cpp:74-147 */
  g_feas_sections++;
  if (t2 - t1 < minT) return RQCD_INPUT_INDETERMINABLE;
  if (dmax(thrust(f, t1), thrust(f, t2)) > fmaxA) return RQCD_INPUT_INFEASIBLE_THRUST_HIGH;
  if (dmin(thrust(f, t1), thrust(f, t2)) < fminA) return RQCD_INPUT_INFEASIBLE_THRUST_LOW;
  double fminSqr = 0, fmaxSqr = 0, jmaxSqr = 0;
  for (int i = 0; i < 3; i++) {
    double amin, amax;
    g_feas_axis_iters++;
    minmax_acc(&f->ax[i], f->peak[i], &f->peak_init[i], &amin, &amax, t1, t2);
    double v1 = amin - f->grav[i];
    double v2 = amax - f->grav[i];
    if (dmax(sq(v1), sq(v2)) > sq(fmaxA)) return RQCD_INPUT_INFEASIBLE_THRUST_HIGH;
    if (v1 * v2 < 0) {
      fminSqr += 0;
    } else {
      fminSqr += sq(dmin(fabs(v1), fabs(v2)));
    }
    fmaxSqr += sq(dmax(fabs(v1), fabs(v2)));
    jmaxSqr += max_jerk_sq(&f->ax[i], t1, t2);
  }
  g_feas_full++;
  double fmin = sqrt(fminSqr);
  double fmax = sqrt(fmaxSqr);
  double wBound;
  if (fminSqr > 1e-6)
    wBound = sqrt(jmaxSqr / fminSqr);
  else
    wBound = 1.7976931348623157e308; /* numeric_limits<double>::max() */
  if (fmax < fminA) return RQCD_INPUT_INFEASIBLE_THRUST_LOW;
  if (fmin > fmaxA) return RQCD_INPUT_INFEASIBLE_THRUST_HIGH;
  if (fmin < fminA || fmax > fmaxA || wBound > wmaxA) {
    double tHalf = (t1 + t2) / 2;
    int r1 = feas_section(f, fminA, fmaxA, wmaxA, t1, tHalf, minT);
    if (r1 == RQCD_INPUT_FEASIBLE) return feas_section(f, fminA, fmaxA, wmaxA, tHalf, t2, minT);
    return r1;
  }
  return RQCD_INPUT_FEASIBLE;
}

void orc_input_feasibility(const rqcd_bc_soa* in, int64_t n, const double gravity[3], double fmin, double fmax,
                           double wmax, double min_time_section, const int64_t* group, uint8_t* result) {
  feas_ctx f;
  memset(&f, 0, sizeof f);
  for (int64_t i = 0; i < n; i++) {
    double coef[18], cost, Tf;
    /* a new generator object (constructor -> SetInitialState -> Reset) whenever the group changes */
    if (!group || i == 0 || group[i] != group[i - 1]) f.peak_init[0] = f.peak_init[1] = f.peak_init[2] = 0;
    gen_traj(in, i, f.ax, coef, &cost, &Tf);
    for (int k = 0; k < 3; k++) f.grav[k] = gravity[k];
    result[i] = (uint8_t)feas_section(&f, fmin, fmax, wmax, 0, Tf, min_time_section);
  }
}

/* ------------------------------------------------------------------------------------------------
 * Synthetic inputs
 * ---------------------------------------------------------------------------------------------- */
static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

/* u in [0,1) from (seed, global index, row); then lo + u*(hi-lo).  Ranges: PerformanceEval.cpp:59-65. */
static inline double synth_u(uint64_t seed, uint64_t idx, uint64_t row) {
  uint64_t h = splitmix64(splitmix64(seed ^ (row * 0xD6E8FEB86659FD93ull)) + idx);
  return (double)(h >> 11) * 0x1.0p-53;
}

void orc_synth_bc(uint64_t seed, int64_t index_base, int64_t n, double* bc, int64_t stride) {
  for (int64_t i = 0; i < n; i++) {
    uint64_t g = (uint64_t)(index_base + i);
    for (int r = 0; r < RQCD_BC_ROWS; r++) {
      double u = synth_u(seed, g, (uint64_t)r);
      double v;
      if (r < RQCD_BC_V0)
        v = 0.0; /* pos0 = 0 */
      else if (r == RQCD_BC_TF)
        v = 0.2 + u * (4.0 - 0.2);
      else
        v = -4.0 + u * 8.0;
      bc[r * stride + i] = v;
    }
  }
}

/* mt19937 (std::mt19937, seed 0 -> init_genrand) */
typedef struct mt_t {
  uint32_t s[624];
  int idx;
} mt_t;
static void mt_seed(mt_t* m, uint32_t seed) {
  m->s[0] = seed;
  for (int i = 1; i < 624; i++) m->s[i] = 1812433253u * (m->s[i - 1] ^ (m->s[i - 1] >> 30)) + (uint32_t)i;
  m->idx = 624;
}
static uint32_t mt_next(mt_t* m) {
  if (m->idx >= 624) {
    for (int i = 0; i < 624; i++) {
      uint32_t y = (m->s[i] & 0x80000000u) | (m->s[(i + 1) % 624] & 0x7fffffffu);
      m->s[i] = m->s[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    m->idx = 0;
  }
  uint32_t y = m->s[m->idx++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}
/* libstdc++ generate_canonical<double,53> with a 32-bit engine: two draws, low word first. */
static double mt_canonical(mt_t* m) {
  double sum = 0, tmp = 1;
  sum += (double)mt_next(m) * tmp;
  tmp *= 4294967296.0;
  sum += (double)mt_next(m) * tmp;
  tmp *= 4294967296.0;
  double ret = sum / tmp;
  if (ret >= 1.0) ret = nextafter(1.0, 0.0);
  return ret;
}
static double mt_uniform(mt_t* m, double a, double b) { return mt_canonical(m) * (b - a) + a; }

/* g++ evaluates the three arguments of `Vec3 v(r(gen), r(gen), r(gen))` right to left on x86-64, so the
 * first draw lands in z (checked against oracle/_ref and the golden CSV in tests/test_oracle_golden.py). */
static void draw3(mt_t* m, double a, double b, double out[3]) {
  out[2] = mt_uniform(m, a, b);
  out[1] = mt_uniform(m, a, b);
  out[0] = mt_uniform(m, a, b);
}

int64_t orc_c1_inputs(int64_t n_accept, int filter, double* bc, int64_t stride, double* spheres,
                      int64_t sphere_stride) {
  mt_t m;
  mt_seed(&m, 0);
  const double grav[3] = {0, 0, -9.81};
  int64_t trials = 0, i = 0;
  while (i < n_accept) { /* PerformanceEval.cpp:117-173 */
    trials++;
    double v0[3], a0[3], pf[3], vf[3], af[3], op[3];
    draw3(&m, -4, 4, v0);
    draw3(&m, -4, 4, a0);
    draw3(&m, -4, 4, pf);
    draw3(&m, -4, 4, vf);
    draw3(&m, -4, 4, af);
    double Tf = mt_uniform(&m, 0.2, 4.0);
    draw3(&m, -4, 4, op);
    double rad = mt_uniform(&m, 0.1, 1.5);
    for (int k = 0; k < 3; k++) {
      bc[(RQCD_BC_P0 + k) * stride + i] = 0;
      bc[(RQCD_BC_V0 + k) * stride + i] = v0[k];
      bc[(RQCD_BC_A0 + k) * stride + i] = a0[k];
      bc[(RQCD_BC_PF + k) * stride + i] = pf[k];
      bc[(RQCD_BC_VF + k) * stride + i] = vf[k];
      bc[(RQCD_BC_AF + k) * stride + i] = af[k];
      spheres[k * sphere_stride + i] = op[k];
    }
    bc[RQCD_BC_TF * stride + i] = Tf;
    spheres[3 * sphere_stride + i] = rad;
    if (filter) {
      rqcd_bc_soa one = {bc + i, stride, RQCD_GOAL_ALL, 0};
      uint8_t res;
      orc_input_feasibility(&one, 1, grav, 5, 30, 20, 0.002, NULL, &res);
      if (res != RQCD_INPUT_FEASIBLE) continue;
    }
    i++;
  }
  return trials;
}

void orc_c2_inputs(int64_t n_batches, int64_t per_batch, double* bc, int64_t stride) {
  mt_t m;
  mt_seed(&m, 0);
  int64_t i = 0;
  for (int64_t b = 0; b < n_batches; b++) { /* ForestPerformanceEval.cpp:167-193 */
    double v0[3], a0[3];
    /* Vec3 vel0(randVelX(gen), randVelYZ(gen), randVelYZ(gen)): right-to-left, z drawn first */
    v0[2] = mt_uniform(&m, -2, 2);
    v0[1] = mt_uniform(&m, -2, 2);
    v0[0] = mt_uniform(&m, 2, 8);
    a0[2] = mt_uniform(&m, -2, 2);
    a0[1] = mt_uniform(&m, -2, 2);
    a0[0] = mt_uniform(&m, 4, 10);
    for (int64_t t = 0; t < per_batch; t++, i++) {
      double pf[3];
      draw3(&m, -2.5, 2.5, pf);
      double Tf = mt_uniform(&m, 0.5, 2.0);
      bc[(RQCD_BC_P0 + 0) * stride + i] = -2.5;
      bc[(RQCD_BC_P0 + 1) * stride + i] = 0;
      bc[(RQCD_BC_P0 + 2) * stride + i] = 0;
      for (int k = 0; k < 3; k++) {
        bc[(RQCD_BC_V0 + k) * stride + i] = v0[k];
        bc[(RQCD_BC_A0 + k) * stride + i] = a0[k];
        bc[(RQCD_BC_PF + k) * stride + i] = pf[k];
        bc[(RQCD_BC_VF + k) * stride + i] = 0;
        bc[(RQCD_BC_AF + k) * stride + i] = 0;
      }
      bc[RQCD_BC_TF * stride + i] = Tf;
    }
  }
}

void orc_forest_obstacles(rqcd_obstacle out[5]) { /* ForestPerformanceEval.cpp:130-136 */
  static const double pos[5][3] = {{-1.75, 1.5, 0}, {0.5, -1.5, 0}, {1.5, 0.5, 0}, {-1.0, -1.0, 0}, {0.0, 0.8, -0.3}};
  memset(out, 0, sizeof(rqcd_obstacle) * 5);
  for (int i = 0; i < 5; i++) {
    out[i].type = RQCD_RECT_PRISM;
    for (int k = 0; k < 3; k++) out[i].center[k] = pos[i][k];
    out[i].side[0] = 0.5;
    out[i].side[1] = 0.5;
    out[i].side[2] = 5;
    out[i].quat[0] = 1; /* Rotation::Identity, Rotation.hpp:59-61 */
  }
  /* Rotation::FromAxisAngle(Vec3(1,0,0).GetUnitVector(), +-45*M_PI/180) -- Rotation.hpp:92-97 */
  for (int i = 3; i < 5; i++) {
    double angle = (i == 3 ? 45 : -45) * M_PI / 180;
    v3 u = v3_unit((v3){1, 0, 0});
    out[i].quat[0] = cos(angle * 0.5);
    out[i].quat[1] = sin(angle * 0.5) * u.x;
    out[i].quat[2] = sin(angle * 0.5) * u.y;
    out[i].quat[3] = sin(angle * 0.5) * u.z;
  }
}
