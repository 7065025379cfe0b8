// tools/mathcheck.cu -- accuracy of the device transcendental helpers used by solve_p3, against host long double.
// Build twice (with -fmad=false and default) to see whether the flag changes the CUDA math library's accuracy.
#include <cstdio>
#include <cmath>
#include <vector>
#include <random>
#include "../rapidquadcoptercollisiondetection_b200/csrc/rqcd_device.cuh"
using namespace rqcd;
__global__ void k(const double* x, const double* u, int n, double* o_cbrt, double* o_acos, double* o_s, double* o_c, double* o_mycbrt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  o_cbrt[i] = cbrt(x[i]);
  o_acos[i] = acos_unit(u[i]);
  double s, c;
  sincos_third(acos_unit(u[i]) / 3, s, c);
  o_s[i] = s; o_c[i] = c;
#ifdef HAVE_MYCBRT
  o_mycbrt[i] = cbrt_pos(x[i]);
#else
  o_mycbrt[i] = cbrt(x[i]);
#endif
}
static double ulp_err(double got, long double want) {
  if (want == 0) return got == 0 ? 0 : 1e300;
  int e; frexp((double)want, &e);
  long double ulp = ldexpl(1.0L, e - 53);
  return (double)(fabsl((long double)got - want) / ulp);
}
int main() {
  const int n = 1 << 20;
  std::mt19937_64 g(1);
  std::vector<double> x(n), u(n);
  std::uniform_real_distribution<double> le(-40, 40), uu(-1, 1);
  for (int i = 0; i < n; i++) { x[i] = exp(le(g)); u[i] = uu(g); if (i % 17 == 0) u[i] = copysign(1 - exp(le(g) - 41), u[i]); }
  double *dx, *du, *o[5];
  cudaMalloc(&dx, n * 8); cudaMalloc(&du, n * 8);
  for (auto& p : o) cudaMalloc(&p, n * 8);
  cudaMemcpy(dx, x.data(), n * 8, cudaMemcpyHostToDevice); cudaMemcpy(du, u.data(), n * 8, cudaMemcpyHostToDevice);
  k<<<n / 256, 256>>>(dx, du, n, o[0], o[1], o[2], o[3], o[4]);
  std::vector<double> r[5];
  for (int j = 0; j < 5; j++) { r[j].resize(n); cudaMemcpy(r[j].data(), o[j], n * 8, cudaMemcpyDeviceToHost); }
  double m[5] = {0, 0, 0, 0, 0};
  for (int i = 0; i < n; i++) {
    m[0] = fmax(m[0], ulp_err(r[0][i], cbrtl((long double)x[i])));
    long double ac = acosl((long double)u[i]);
    m[1] = fmax(m[1], ulp_err(r[1][i], ac));
    long double th = (long double)(r[1][i] / 3);  // the device's own argument
    m[2] = fmax(m[2], ulp_err(r[2][i], sinl(th)));
    m[3] = fmax(m[3], ulp_err(r[3][i], cosl(th)));
    m[4] = fmax(m[4], ulp_err(r[4][i], cbrtl((long double)x[i])));
  }
  printf("max ulp error: cbrt %.2f acos_unit %.2f sin_third %.2f cos_third %.2f cbrt_pos %.2f  (%s)\n", m[0], m[1], m[2], m[3], m[4], cudaGetErrorString(cudaGetLastError()));
  return 0;
}
