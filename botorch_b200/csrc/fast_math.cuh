// FP64 exp / log / log1p for the utility kernels (sample_reduce*, log_hvi, cov).
//
// Why not libdevice: its polynomials carry their coefficients as 64-bit immediates, which sm_100a materialises with two UMOV
// per coefficient in front of every DFMA -- 22 UMOV per exp or log call, 36 per log1p.  In the transcendental-bound kernels
// that is 12-22 % of all issued instructions (profiles/r02_simt_instruction_mix.md), and the FP64 pipe idles while they
// issue.  Here the coefficients live in constant memory and reach the DFMAs through uniform registers, two per LDCU.128, and
// the special-case handling shrinks to one range test with the libdevice function as the out-of-range fallback:
//   exp: 16 FP64 + 5 LDCU + ~8 integer instructions (libdevice: ~60 executed), log: ~30 FP64 + 5 LDCU + ~10 (~85).
// Accuracy: <= 2 ulp over the fast ranges, checked against long double on the host (tools/fast_math/check.cpp, the same
// source compiled for the CPU) and against torch on the device (tests/test_gpu_fast_math.py).  Coefficients: Chebyshev fits
// at 60 digits, tools/fast_math/gen_coeffs.py.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define FM_HD __device__ __forceinline__
#define FM_CONST static __constant__
#else
#define FM_HD inline
#define FM_CONST static const
#endif

namespace mcacq {

//   exp(r) = 1 + r + r^2 Q(r), |r| <= ln2 / 2: Q lowest power first (fit error 1.1e-16 of Q, i.e. < 1.3e-17 of exp)
FM_CONST double FM_EXP_Q[10] = {0.50000000000000011, 0.16666666666666669, 0.04166666666662399, 0.0083333333333300511,
                                0.0013888888917281794, 0.00019841269863105968, 2.4801521190217729e-05,
                                2.7557268378684192e-06, 2.7620138719733994e-07, 2.5100424157005067e-08};
//   log(m) = 2 f + 2 f s G(s), f = (m - 1) / (m + 1), s = f^2, m in [sqrt(1/2), sqrt 2]: G lowest power first (1.0e-18)
FM_CONST double FM_LOG_G[8] = {0.33333333333333331, 0.20000000000000442, 0.14285714285399892, 0.11111111196290523,
                               0.090908977737605687, 0.076931223688108327, 0.066343191846334124, 0.065440174254637773};
//   ln 2 = hi + lo, hi with its low 21 mantissa bits clear (e * hi is exact for |e| < 2^11)
FM_CONST double FM_LN2[4] = {0.69314718036912382, 1.9082149292705877e-10, 1.4426950408889634, 6755399441055744.0};

FM_HD int fm_hi(double x) {
#ifdef __CUDA_ARCH__
  return __double2hiint(x);
#else
  uint64_t u; memcpy(&u, &x, 8); return (int)(u >> 32);
#endif
}
FM_HD int fm_lo(double x) {
#ifdef __CUDA_ARCH__
  return __double2loint(x);
#else
  uint64_t u; memcpy(&u, &x, 8); return (int)(u & 0xffffffffu);
#endif
}
FM_HD double fm_make(int hi, int lo) {
#ifdef __CUDA_ARCH__
  return __hiloint2double(hi, lo);
#else
  uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double x; memcpy(&x, &u, 8); return x;
#endif
}
// 1 / x for normal x of moderate magnitude: hardware seed (~20 bits) + two Newton steps
FM_HD double fm_rcp(double x) {
  double r;
#ifdef __CUDA_ARCH__
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#else
  r = 1.0 / x;                     // host model of the ~20-bit seed: the quotient with its low 32 mantissa bits cleared
  { uint64_t u; memcpy(&u, &r, 8); u &= 0xffffffff00000000ull; memcpy(&r, &u, 8); }
#endif
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

// out-of-range fallbacks: libdevice, NOT inlined (rare paths; inlining them would double the code of every call site)
#ifdef __CUDACC__
#define FM_SLOW static __device__ __noinline__
#else
#define FM_SLOW inline
#endif
FM_SLOW double fm_exp_slow(double x) { return exp(x); }
FM_SLOW double fm_log_slow(double x) { return log(x); }
FM_SLOW double fm_log1p_slow(double x) { return log1p(x); }

// exp(x).  Fast range |x| <= 708 (the result and its scaling stay normal); libdevice outside (overflow, denormal results, NaN).
FM_HD double fm_exp(double x) {
  if (!(fabs(x) <= 708.0)) return fm_exp_slow(x);
  const double t = fma(x, FM_LN2[2], FM_LN2[3]);     // round(x / ln2) in the low mantissa bits of t
  const int k = fm_lo(t);
  const double kf = t - FM_LN2[3];
  double r = fma(kf, -FM_LN2[0], x);
  r = fma(kf, -FM_LN2[1], r);
  double q = FM_EXP_Q[9];
#pragma unroll
  for (int i = 8; i >= 0; i--) q = fma(q, r, FM_EXP_Q[i]);
  const double p = fma(r * r, q, r) + 1.0;            // in (0.70, 1.42)
  return fm_make(fm_hi(p) + (k << 20), fm_lo(p));     // p * 2^k, k in [-1022, 1022]
}

// log(x).  Fast range: normal positive finite x; libdevice otherwise (0, denormals, negatives, inf, NaN).
FM_HD double fm_log(double x) {
  int hi = fm_hi(x);
  if (!(hi >= 0x00100000 && hi < 0x7ff00000)) return fm_log_slow(x);
  int e = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;                // mantissa in [1, 2)
  if (hi >= 0x3ff6a09f) { hi -= 0x00100000; e += 1; } // -> [sqrt(1/2), sqrt 2) (to within the cut's 32-bit granularity)
  const double m = fm_make(hi, fm_lo(x));
  const double f = m - 1.0;                           // exact
  const double g = m + 1.0;
  const double rg = fm_rcp(g);
  double q = f * rg;
  q = fma(fma(-q, g, f), rg, q);                      // f / g to ~1 ulp
  const double s = q * q;
  double p = FM_LOG_G[7];
#pragma unroll
  for (int i = 6; i >= 0; i--) p = fma(p, s, FM_LOG_G[i]);
  const double ef = (double)e;
  const double q2 = q + q;
  const double tail = fma(ef, FM_LN2[1], q2 * s * p);  // e * ln2_lo + 2 q s G(s)
  return fma(ef, FM_LN2[0], q2 + tail);
}

// log1p(x) for x >= 0 (softplus: x = exp(u)): log(1 + x) plus the rounding error of the sum, first order
FM_HD double fm_log1p_nonneg(double x) {
  if (!(x >= 0.0 && x < 1e300)) return fm_log1p_slow(x);
  const double u = 1.0 + x;
  const double c = x - (u - 1.0);                     // exact: what the sum lost
  return fm_log(u) + c * fm_rcp(u);
}

}  // namespace mcacq
