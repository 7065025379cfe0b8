// tools/divcheck.cu -- the regrouped divisions and square roots of rqcd_device.cuh (FastMath: shared reciprocal
// refinement, accumulated range tests) against the compiler's own FP64 `/` and sqrt(), bit for bit, over random operands of every exponent range
// plus zeros, subnormals, infinities and NaNs. Prints the number of differing quotients (must be 0).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/divcheck tools/divcheck.cu && tools/divcheck [rounds]
#include <cstdio>
#include <cstdlib>
#include "../rapidquadcoptercollisiondetection_b200/csrc/rqcd_device.cuh"
using namespace rqcd;

__device__ unsigned long long sm64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
// a double with random sign and mantissa; exponent class chosen by `mode`
__device__ double rnd(unsigned long long h, int mode) {
  const unsigned long long mant = h & 0x000fffffffffffffull, sign = h & 0x8000000000000000ull;
  unsigned long long e;
  const unsigned pick = (unsigned)(h >> 52) & 0x7ff;
  switch (mode) {
    case 0: e = 1023 - 8 + pick % 17; break;            // around 1
    case 1: e = 1023 - 60 + pick % 121; break;          // moderate
    case 2: e = pick; break;                            // anything, incl. 0 (subnormal/zero) and 0x7ff (inf/nan)
    case 3: e = pick % 80; break;                       // tiny
    default: e = 0x7ff - pick % 80; break;              // huge
  }
  unsigned long long m = mant;
  if ((h >> 40 & 0xff) == 0) m = 0;                     // exact powers of two / zero / inf now and then
  return __longlong_as_double((long long)(sign | (e << 52) | m));
}
__device__ bool same(double a, double b) {
  return __double_as_longlong(a) == __double_as_longlong(b) || (a != a && b != b);
}
__global__ void k(unsigned long long seed, unsigned long long* bad, unsigned long long* slow) {
  const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  unsigned long long nbad = 0, nslow = 0;
  for (int it = 0; it < 64; it++) {
    const unsigned long long h = sm64(seed + i * 64 + it);
    const int my = (int)(h % 5), mx = (int)((h / 5) % 5);
    const double y = rnd(sm64(h ^ 1), my);
    const double x0 = rnd(sm64(h ^ 2), mx), x2 = rnd(sm64(h ^ 4), (mx + 1) % 5);
    double x1 = rnd(sm64(h ^ 3), mx);
    if (it % 16 == 5) x1 = copysign(0.0, x1);  // signed zero numerators
    {  // three quotients that share a divisor; half of them through the zero-aware form
      FastMath fm;
      const FastMath::Div d = fm.divisor(y);
      const double q0 = fm.div(x0, d), q1 = fm.div_z(x1, d), q2 = (it & 1) ? fm.div(x2, d) : fm.div_z(x2, d);
      if (fm.ok) nbad += !same(q0, x0 / y) + !same(q1, x1 / y) + !same(q2, x2 / y);
      else nslow++;
    }
    {  // square roots: plain, zero-aware, and of arguments that must be flagged (negative, NaN, tiny, huge)
      FastMath fa, fb, fc;
      const double a0 = fabs(x0), s0 = fa.sqrt(a0), s1 = fb.sqrt_z((it & 3) ? fabs(x1) : 0.0), s2 = fc.sqrt(x2);
      if (fa.ok) nbad += !same(s0, sqrt(a0));
      if (fb.ok) nbad += !same(s1, sqrt((it & 3) ? fabs(x1) : 0.0));
      if (fc.ok) nbad += !same(s2, sqrt(x2));
    }
    {  // the constant divisors of solveP3
      FastMath f3, f9, f54;
      const double c0 = f3.div(x0, f3.divisor(3.0, kRcpConst[0])), c1 = f9.div_z(x1, f9.divisor(9.0, kRcpConst[1]));
      const double c2 = f54.div(x2, f54.divisor(54.0, kRcpConst[2]));
      if (f3.ok) nbad += !same(c0, x0 / 3);
      if (f9.ok) nbad += !same(c1, x1 / 9);
      if (f54.ok) nbad += !same(c2, x2 / 54);
    }
  }
  atomicAdd(bad, nbad);
  atomicAdd(slow, nslow);
}
__global__ void k_init(double* out) {
  out[0] = rcp_refined(3.0), out[1] = rcp_refined(9.0), out[2] = rcp_refined(54.0);
}
int main(int argc, char** argv) {
  const int rounds = argc > 1 ? atoi(argv[1]) : 64;
  double* rc;
  cudaMalloc(&rc, 24);
  k_init<<<1, 1>>>(rc);
  cudaMemcpyToSymbol(kRcpConst, rc, 24, 0, cudaMemcpyDeviceToDevice);
  unsigned long long *d, h[2] = {0, 0};
  cudaMalloc(&d, 16);
  cudaMemset(d, 0, 16);
  const int blocks = 148 * 8, threads = 256;
  for (int r = 0; r < rounds; r++) k<<<blocks, threads>>>(0x1234567ull + (unsigned long long)r * blocks * threads * 64, d, d + 1);
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  const double groups = (double)rounds * blocks * threads * 64;
  printf("divcheck: %.3e divisor groups (3 shared-divisor + 3 constant-divisor quotients + 3 square roots each), %llu differing results, "
         "%.2f%% of the groups left the fast range  (%s)\n",
         groups, h[0], 100.0 * h[1] / groups, cudaGetErrorString(cudaGetLastError()));
  return h[0] != 0;
}
