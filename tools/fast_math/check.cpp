// Host check of botorch_b200/csrc/fast_math.cuh (same source, compiled for the CPU; the reciprocal seed is modelled by a
// 20-bit truncation): maximum error in ulp against long double over dense random samples of the ranges the kernels use.
//   g++ -O2 -std=c++17 -o check check.cpp && ./check
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "../../botorch_b200/csrc/fast_math.cuh"

static double ulp_err(double got, long double want) {
  if (want == 0.0L) return got == 0.0 ? 0.0 : 1e9;
  int e;
  std::frexp((double)want, &e);
  const long double ulp = std::ldexp(1.0L, e - 53);
  return (double)(fabsl((long double)got - want) / ulp);
}

int main(int argc, char** argv) {
  std::mt19937_64 rng(1);
  auto uni = [&](double a, double b) { return a + (b - a) * (double)(rng() >> 11) * (1.0 / 9007199254740992.0); };
  const int N = argc > 1 ? atoi(argv[1]) : 20000000;   // samples per range
  double worst;
  // exp on [-708, 708] and dense near 0
  worst = 0;
  for (int i = 0; i < N; i++) {
    const double x = (i & 1) ? uni(-708, 708) : uni(-2, 2) * std::pow(10.0, -uni(0, 12));
    worst = std::fmax(worst, ulp_err(mcacq::fm_exp(x), expl((long double)x)));
  }
  printf("fm_exp          max %.3f ulp\n", worst);
  // log over the whole normal range (log-uniform), near 1, and on [1, 32] (fatmin sums)
  worst = 0;
  double worst1 = 0, worst32 = 0;
  for (int i = 0; i < N; i++) {
    const double a = std::ldexp(uni(1, 2), (int)uni(-1022, 1023));
    worst = std::fmax(worst, ulp_err(mcacq::fm_log(a), logl((long double)a)));
    const double b = 1.0 + uni(-1, 1) * std::pow(10.0, -uni(0.31, 15));
    worst1 = std::fmax(worst1, ulp_err(mcacq::fm_log(b), logl((long double)b)));
    const double c = uni(1, 32);
    worst32 = std::fmax(worst32, ulp_err(mcacq::fm_log(c), logl((long double)c)));
  }
  printf("fm_log          max %.3f ulp (normal range), %.3f (near 1), %.3f ([1, 32])\n", worst, worst1, worst32);
  // log1p on [0, e^20] log-uniform down to 1e-320
  worst = 0;
  for (int i = 0; i < N; i++) {
    const double x = std::exp(uni(-737, 20.0));
    worst = std::fmax(worst, ulp_err(mcacq::fm_log1p_nonneg(x), log1pl((long double)x)));
  }
  printf("fm_log1p_nonneg max %.3f ulp\n", worst);
  // special values take the libdevice / libm fallback
  printf("specials: exp(-800)=%g exp(800)=%g exp(nan)=%g log(0)=%g log(-1)=%g log(inf)=%g log(5e-324)=%g log1p(0)=%g\n", mcacq::fm_exp(-800),
         mcacq::fm_exp(800), mcacq::fm_exp(NAN), mcacq::fm_log(0.0), mcacq::fm_log(-1.0), mcacq::fm_log(INFINITY), mcacq::fm_log(5e-324),
         mcacq::fm_log1p_nonneg(0.0));
  return 0;
}
