// Fused log-hypervolume-improvement of one MC sample: the whole inclusion-exclusion loop of qLogEHVI / qLogNEHVI
// (`_compute_log_qehvi`, botorch/acquisition/multi_objective/logei.py:272-435, steps 1-8) in one kernel.
//
//   out[B] = logsumexp_c logdiffexp( E_c, O_c ),        E_c / O_c = logsumexp over the subsets s of {0..q-1} with even / odd |s| of
//   area(c, s) = sum_k fatmin2( fatmin_{j in s} li[j][k], log_cell_length[c][k] ),   li[j][k] = log_fatplus(obj[j][k] - cell_lower[c][k])
//
// The reference (and the per-subset-size kernels of log_areas.cu that port its `logei_fused.cpp`) evaluate li[j][k] once per
// SUBSET MEMBERSHIP -- sum_i i C(q, i) = q 2^(q-1) times per cell (32 for q = 4) instead of q times -- and leave the subset /
// cell reductions to ~100 small tensor ops over (B, nc, n_sub) intermediates.  Here one thread owns one sample: it keeps
// obj[q][m] and, per cell, li[q][m] in registers, enumerates the subsets as bit masks (members in increasing index order,
// like itertools.combinations), and folds areas -> parity sums -> cells with streaming log-sum-exps.  Nothing but obj, the
// cell bounds and out[B] touches memory.  Arithmetic of the building blocks (log_fatplus with the safe softplus, fatmin with
// its n == 1 and < -1e29 branches, the 1e10 clamp of the upper bound) is shared with log_areas.cu, i.e. follows
// botorch/csrc/logei_fused.cpp:39-177, 218-238; the log-space reductions follow utils/safe_math.py:36-46, 100-143.
//
// Backward: the same thread evaluates the cell's subset areas once (kept in registers for q <= 4), turns them into the
// weights w_c * d(logdiffexp)/d(E|O) * exp(area - E|O), and accumulates d out / d obj[j][k] one objective at a time.
#include <math_constants.h>

#include <cstdlib>

#include "common.cuh"

namespace mcacq {

// 1 / x for normal-range x: hardware seed + two Newton steps (error < 1 ulp-ish; the denominators here are >= 1)
__device__ __forceinline__ double hv_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

__device__ __forceinline__ double hv_softplus(double y) {
  if (y > 20.0) return y;
  if (y < -20.0) return fm_exp(y);
  return fm_log1p_nonneg(fm_exp(y));
}
__device__ __forceinline__ double hv_sigmoid(double y) {
  if (y >= 0.0) { const double e = fm_exp(-y); return 1.0 / (1.0 + e); }
  const double e = fm_exp(y);
  return e / (1.0 + e);
}
template <bool GRAD>
__device__ __forceinline__ double hv_log_fatplus(double x, double tau, double inv_tau, double& grad) {
  const double y = x * inv_tau;
  const double den = fma(y, y, 1.0);
  const double cy = (den < 1e300) ? hv_rcp(den) : 0.0;
  const double f = hv_softplus(y) + 0.1 * cy;
  const double tf = tau * f;
  if (!(tf > 0.0)) { if (GRAD) grad = 0.0; return -1e30; }
  if (GRAD) grad = (hv_sigmoid(y) - 0.2 * y * cy * cy) / tf;
  return fm_log(tf);
}

// fatmin over the members of `mask` of column k of li (QMAX x MMAX, row-major in registers); gw[j] = d / d li[j][k].
// `mask` is a compile-time constant wherever the caller's subset loop is unrolled (q <= 4), so the membership tests fold.
template <int QMAX, int MMAX, bool GRAD>
__device__ __forceinline__ double hv_fatmin_masked(const double (&li)[QMAX * MMAX], int k, unsigned mask, int n, double tau,
                                                   double inv_tau, double (&gw)[QMAX]) {
  double mn = CUDART_INF;
  int ami = -1;
#pragma unroll
  for (int j = 0; j < QMAX; j++)
    if ((mask >> j) & 1u) { const double v = li[j * MMAX + k]; if (ami < 0 || v < mn) { mn = v; ami = j; } }
  if (n == 1 || mn < -1e29) {
    if (GRAD) {
#pragma unroll
      for (int j = 0; j < QMAX; j++) gw[j] = (j == ami) ? 1.0 : 0.0;
    }
    return mn;
  }
  double S = 0.0, S_pd = 0.0;
  double pd[QMAX];
#pragma unroll
  for (int j = 0; j < QMAX; j++) {
    pd[j] = 0.0;
    if ((mask >> j) & 1u) {
      const double z = (li[j * MMAX + k] - mn) * inv_tau;
      const double d = fma(z, z + 2.0, 2.0);
      const double rd = (d < 1e300) ? hv_rcp(d) : 0.0;
      S += 2.0 * rd;
      if (GRAD) { pd[j] = -2.0 * (2.0 + 2.0 * z) * rd * rd; S_pd += pd[j]; }
    }
  }
  if (GRAD) {
    const double rS = hv_rcp(S);
#pragma unroll
    for (int j = 0; j < QMAX; j++)
      gw[j] = ((mask >> j) & 1u) ? ((j == ami) ? 1.0 + (S_pd + 1.0) * rS : -pd[j] * rS) : 0.0;
  }
  return mn - tau * fm_log(S);
}

// fatmin of the pair (a, b); g0 = d / d a.  The smaller entry contributes exactly 2 / 2 = 1 to the sum.
template <bool GRAD>
__device__ __forceinline__ double hv_fatmin2(double a, double b, double tau, double inv_tau, double& g0) {
  const bool a_min = !(b < a);   // first index wins ties, like the strict `<` scan of the n-ary version
  const double mn = a_min ? a : b;
  if (mn < -1e29) { if (GRAD) g0 = a_min ? 1.0 : 0.0; return mn; }
  const double z = ((a_min ? b : a) - mn) * inv_tau;     // the other entry's offset (>= 0)
  const double d = fma(z, z + 2.0, 2.0);
  const double rd = (d < 1e300) ? hv_rcp(d) : 0.0;
  const double S = 1.0 + 2.0 * rd;
  if (GRAD) {
    const double p_other = -2.0 * (2.0 + 2.0 * z) * rd * rd, p_min = -1.0;   // pd of the minimum: -2 * 2 / 4
    const double rS = hv_rcp(S);
    g0 = a_min ? 1.0 + (p_min + p_other + 1.0) * rS : -p_other * rS;
  }
  return mn - tau * fm_log(S);
}

// streaming log-sum-exp with the inf conventions of safe_math.logsumexp (an infinite maximum wins; -inf terms vanish)
struct HvLse {
  double m, s;
  __device__ __forceinline__ void init() { m = -CUDART_INF; s = 0.0; }
  __device__ __forceinline__ void push(double v) {
    if (v == -CUDART_INF) return;
    if (v > m) { s = (m == -CUDART_INF) ? 1.0 : s * fm_exp(m - v) + 1.0; m = v; }
    else s += fm_exp(v - m);
  }
  __device__ __forceinline__ double value() const { return (m == -CUDART_INF) ? m : (isinf(m) ? m : m + fm_log(s)); }
};

// logdiffexp(log_a = E, log_b = O) = O + log1mexp(E - O)   (safe_math.py:106-120, 36-46); dE, dO: partial derivatives
template <bool GRAD>
__device__ __forceinline__ double hv_logdiffexp(double E, double O, double& dE, double& dO) {
  if (O == -CUDART_INF) { if (GRAD) { dE = 0.0; dO = 0.0; } return -CUDART_INF; }
  const double x = E - O;
  const double ex = exp(x);                       // E = -inf -> 0
  const double l1m = (-0.69314718055994530942 < x) ? log(-expm1(x)) : log1p(-ex);
  if (GRAD) { const double fp = -ex / (1.0 - ex); dE = fp; dO = 1.0 - fp; }
  return O + l1m;
}

template <int QMAX, int MMAX, bool GRAD>
__device__ __forceinline__ double hv_area(const double (&li)[QMAX * MMAX], const double* __restrict__ ll, int m, unsigned mask,
                                          int n, double tau_max, double inv_tm, double (&garea)[QMAX * MMAX]) {
  double area = 0.0;
#pragma unroll
  for (int k = 0; k < MMAX; k++) {
    if (k < m) {
      double gw[QMAX], g0;
      const double lim = hv_fatmin_masked<QMAX, MMAX, GRAD>(li, k, mask, n, tau_max, inv_tm, gw);
      area += hv_fatmin2<GRAD>(lim, ll[k], tau_max, inv_tm, g0);
      if (GRAD) {
#pragma unroll
        for (int j = 0; j < QMAX; j++) garea[j * MMAX + k] = g0 * gw[j];
      }
    }
  }
  return area;
}

// Subset loop: fully unrolled (compile-time masks, the membership tests fold away) up to q = 4, a runtime loop beyond.
template <int QMAX>
struct HvMasks { static constexpr bool UNROLL = QMAX <= 4; static constexpr unsigned NM = 1u << QMAX; };

template <int QMAX, int MMAX, int MINB>
__global__ void __launch_bounds__(128, MINB)
log_hvi_fwd_kernel(const double* __restrict__ obj, const double* __restrict__ cl, const double* __restrict__ lcl, int64_t B, int q,
                   int m, int nc, double tau_relu, double tau_max, double* __restrict__ out) {
  const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double inv_tr = 1.0 / tau_relu, inv_tm = 1.0 / tau_max;
  double o[QMAX * MMAX];
#pragma unroll
  for (int j = 0; j < QMAX; j++)
#pragma unroll
    for (int k = 0; k < MMAX; k++) o[j * MMAX + k] = (j < q && k < m) ? obj[(b * q + j) * m + k] : 0.0;
  HvLse total;
  total.init();
  const unsigned n_masks = 1u << q;
  for (int c = 0; c < nc; c++) {
    const double* lo = cl + (int64_t)c * m;
    const double* ll = lcl + (int64_t)c * m;
    double li[QMAX * MMAX], gdummy[QMAX * MMAX];
#pragma unroll
    for (int j = 0; j < QMAX; j++)
#pragma unroll
      for (int k = 0; k < MMAX; k++) {
        double gd;
        li[j * MMAX + k] = (j < q && k < m) ? hv_log_fatplus<false>(o[j * MMAX + k] - lo[k], tau_relu, inv_tr, gd) : 0.0;
      }
    HvLse par[2];
    par[0].init(); par[1].init();
    if (HvMasks<QMAX>::UNROLL) {
#pragma unroll
      for (unsigned mask = 1; mask < HvMasks<QMAX>::NM; mask++) {
        if (mask < n_masks) {
          const int n = __popc(mask);
          const double area = hv_area<QMAX, MMAX, false>(li, ll, m, mask, n, tau_max, inv_tm, gdummy);
          if (n & 1) par[1].push(area); else par[0].push(area);
        }
      }
    } else {
      for (unsigned mask = 1; mask < n_masks; mask++) {
        const int n = __popc(mask);
        const double area = hv_area<QMAX, MMAX, false>(li, ll, m, mask, n, tau_max, inv_tm, gdummy);
        if (n & 1) par[1].push(area); else par[0].push(area);
      }
    }
    double dE, dO;
    total.push(hv_logdiffexp<false>(par[0].value(), par[1].value(), dE, dO));
  }
  out[b] = total.value();
}

// Backward, q <= 4: per cell (1) all subset areas (kept in registers) and the parity sums, (2) the subset weights
// w_c * d(logdiffexp)/d(E|O) * exp(area - E|O), (3) one objective k at a time, every subset's fatmin gradients weighted into
// d / d li[:, k].  The areas are evaluated once; only the (cheap, rational) fatmin weights are evaluated a second time.
template <int QMAX, int MMAX, int MINB>
__global__ void __launch_bounds__(128, MINB)
log_hvi_bwd_kernel(const double* __restrict__ gout, const double* __restrict__ outv, const double* __restrict__ obj,
                   const double* __restrict__ cl, const double* __restrict__ lcl, int64_t B, int q, int m, int nc,
                   double tau_relu, double tau_max, double* __restrict__ gobj) {
  const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double inv_tr = 1.0 / tau_relu, inv_tm = 1.0 / tau_max;
  double o[QMAX * MMAX], go[QMAX * MMAX];
#pragma unroll
  for (int j = 0; j < QMAX; j++)
#pragma unroll
    for (int k = 0; k < MMAX; k++) {
      o[j * MMAX + k] = (j < q && k < m) ? obj[(b * q + j) * m + k] : 0.0;
      go[j * MMAX + k] = 0.0;
    }
  const double total = outv[b];
  const double g_up = gout[b];
  const unsigned n_masks = 1u << q;
  const bool live = (g_up != 0.0) && !isinf(total) && !isnan(total);
  for (int c = 0; live && c < nc; c++) {
    const double* lo = cl + (int64_t)c * m;
    const double* ll = lcl + (int64_t)c * m;
    double li[QMAX * MMAX], dli[QMAX * MMAX];
#pragma unroll
    for (int j = 0; j < QMAX; j++)
#pragma unroll
      for (int k = 0; k < MMAX; k++) {
        double gd = 0.0;
        li[j * MMAX + k] = (j < q && k < m) ? hv_log_fatplus<true>(o[j * MMAX + k] - lo[k], tau_relu, inv_tr, gd) : 0.0;
        dli[j * MMAX + k] = gd;
      }
    if (HvMasks<QMAX>::UNROLL) {
      constexpr unsigned NM = HvMasks<QMAX>::NM;
      double ws[NM];   // area, then weight, of subset `mask` (index 0 unused)
      double gdummy[QMAX * MMAX];
      HvLse par[2];
      par[0].init(); par[1].init();
#pragma unroll
      for (unsigned mask = 1; mask < NM; mask++) {
        ws[mask] = -CUDART_INF;
        if (mask < n_masks) {
          const int n = __popc(mask);
          ws[mask] = hv_area<QMAX, MMAX, false>(li, ll, m, mask, n, tau_max, inv_tm, gdummy);
          if (n & 1) par[1].push(ws[mask]); else par[0].push(ws[mask]);
        }
      }
      const double E = par[0].value(), O = par[1].value();
      double dE, dO;
      const double dc = hv_logdiffexp<true>(E, O, dE, dO);
      if (dc == -CUDART_INF || isnan(dc)) continue;
      const double wc = g_up * fm_exp(dc - total);          // softmax weight of the cell in the outer log-sum-exp
      if (wc == 0.0) continue;
#pragma unroll
      for (unsigned mask = 1; mask < NM; mask++) {
        const bool odd = __popc(mask) & 1;
        const double ref = odd ? O : E;
        ws[mask] = (mask < n_masks && ws[mask] != -CUDART_INF && ref != -CUDART_INF) ? wc * (odd ? dO : dE) * fm_exp(ws[mask] - ref) : 0.0;
      }
#pragma unroll
      for (int k = 0; k < MMAX; k++) {
        if (k < m) {
          double gl[QMAX];
#pragma unroll
          for (int j = 0; j < QMAX; j++) gl[j] = 0.0;
#pragma unroll
          for (unsigned mask = 1; mask < NM; mask++) {
            if (mask < n_masks) {
              double gw[QMAX], g0;
              const double lim = hv_fatmin_masked<QMAX, MMAX, true>(li, k, mask, __popc(mask), tau_max, inv_tm, gw);
              hv_fatmin2<true>(lim, ll[k], tau_max, inv_tm, g0);
              const double w0 = ws[mask] * g0;
#pragma unroll
              for (int j = 0; j < QMAX; j++)
                if ((mask >> j) & 1u) gl[j] = fma(w0, gw[j], gl[j]);
            }
          }
#pragma unroll
          for (int j = 0; j < QMAX; j++) go[j * MMAX + k] = fma(gl[j], dli[j * MMAX + k], go[j * MMAX + k]);
        }
      }
    } else {
      double gli[QMAX * MMAX], ga[QMAX * MMAX];
#pragma unroll
      for (int i = 0; i < QMAX * MMAX; i++) { gli[i] = 0.0; ga[i] = 0.0; }
      HvLse par[2];
      par[0].init(); par[1].init();
      for (unsigned mask = 1; mask < n_masks; mask++) {
        const int n = __popc(mask);
        const double area = hv_area<QMAX, MMAX, false>(li, ll, m, mask, n, tau_max, inv_tm, ga);
        if (n & 1) par[1].push(area); else par[0].push(area);
      }
      const double E = par[0].value(), O = par[1].value();
      double dE, dO;
      const double dc = hv_logdiffexp<true>(E, O, dE, dO);
      if (dc == -CUDART_INF || isnan(dc)) continue;
      const double wc = g_up * fm_exp(dc - total);
      if (wc == 0.0) continue;
      for (unsigned mask = 1; mask < n_masks; mask++) {
        const int n = __popc(mask);
        const double area = hv_area<QMAX, MMAX, true>(li, ll, m, mask, n, tau_max, inv_tm, ga);
        const double ref = (n & 1) ? O : E;
        if (area == -CUDART_INF || ref == -CUDART_INF) continue;
        const double w = wc * ((n & 1) ? dO : dE) * fm_exp(area - ref);
#pragma unroll
        for (int j = 0; j < QMAX; j++)
#pragma unroll
          for (int k = 0; k < MMAX; k++)
            if (((mask >> j) & 1u) && k < m) gli[j * MMAX + k] = fma(w, ga[j * MMAX + k], gli[j * MMAX + k]);
      }
#pragma unroll
      for (int i = 0; i < QMAX * MMAX; i++) go[i] = fma(gli[i], dli[i], go[i]);
    }
  }
#pragma unroll
  for (int j = 0; j < QMAX; j++)
#pragma unroll
    for (int k = 0; k < MMAX; k++)
      if (j < q && k < m) gobj[(b * q + j) * m + k] = go[j * MMAX + k];
}

__global__ void hv_log_cell_length_kernel(const double* __restrict__ cl, const double* __restrict__ cu, int64_t total,
                                          double* __restrict__ lcl) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < total) lcl[i] = log(min(cu[i], 1e10) - cl[i]);
}

template <int QMAX, int MMAX>
static int log_hvi_launch(int backward, const double* gout, const double* outv, const double* obj, const double* cl,
                          const double* lcl, int64_t B, int q, int m, int nc, double tau_relu, double tau_max, double* out,
                          cudaStream_t st) {
  const unsigned blocks = (unsigned)((B + 127) / 128);
  // resident CTAs per SM the register allocation aims at.  The kernels are latency-bound, so occupancy beats spills: backward
  // (q = m = 4) 2 CTAs = 234 registers, no spills: 45.0 ms per 524288 samples x 32 cells; 3 = 168 registers + 216-byte stack:
  // 32.6 ms; 4 = 128 registers + 368 bytes: 30.8 ms (default); 5 / 6 = 96 / 80 registers: 33.0 / 34.8 ms.  Forward: 112
  // registers at 4 (12.9 ms), spilling at 5 / 6 (14.7 / 15.1 ms).  Only the chosen builds are instantiated (each costs ~40 s of ptxas).
  if (backward)
    log_hvi_bwd_kernel<QMAX, MMAX, 4><<<blocks, 128, 0, st>>>(gout, outv, obj, cl, lcl, B, q, m, nc, tau_relu, tau_max, out);
  else
    log_hvi_fwd_kernel<QMAX, MMAX, 4><<<blocks, 128, 0, st>>>(obj, cl, lcl, B, q, m, nc, tau_relu, tau_max, out);
  count_launch();
  MCACQ_CUDA_CHECK_LAUNCH();
  return 0;
}

static int log_hvi_run(int backward, const double* gout, const double* outv, const double* obj, const double* cl,
                       const double* cu, int64_t B, int q, int m, int nc, double tau_relu, double tau_max, double* out,
                       double* lcl, cudaStream_t st) {
  const int64_t ncell = (int64_t)nc * m;
  hv_log_cell_length_kernel<<<(unsigned)((ncell + 255) / 256), 256, 0, st>>>(cl, cu, ncell, lcl);
  count_launch();
#define HV_CASE(QM, MM) \
  if (q <= QM && m <= MM) return log_hvi_launch<QM, MM>(backward, gout, outv, obj, cl, lcl, B, q, m, nc, tau_relu, tau_max, out, st);
  HV_CASE(2, 2) HV_CASE(4, 2) HV_CASE(2, 4) HV_CASE(4, 4) HV_CASE(6, 2) HV_CASE(6, 4)
#undef HV_CASE
  return MCACQ_ELIMIT;
}

}  // namespace mcacq

static int hv_check(const void* obj, const void* cl, const void* cu, const void* out, const void* lcl, int64_t B, int q, int m,
                    int nc, double tau_relu, double tau_max) {
  if (!obj || !cl || !cu || !out || !lcl || B < 0 || q <= 0 || m <= 0 || nc <= 0) return MCACQ_EINVAL;
  if (!(tau_relu > 0.0) || !(tau_max > 0.0)) return MCACQ_EINVAL;
  if (q > 6 || m > 4) return MCACQ_ELIMIT;
  return 0;
}

extern "C" int mcacq_log_hvi_forward(const double* obj, const double* cell_lower, const double* cell_upper, int64_t B, int q,
                                     int m, int nc, double tau_relu, double tau_max, double* out, double* lcl_workspace,
                                     void* stream) {
  using namespace mcacq;
  int rc = hv_check(obj, cell_lower, cell_upper, out, lcl_workspace, B, q, m, nc, tau_relu, tau_max);
  if (rc) return rc;
  if (B == 0) return 0;
  g_launch_count = 0;
  return log_hvi_run(0, nullptr, nullptr, obj, cell_lower, cell_upper, B, q, m, nc, tau_relu, tau_max, out, lcl_workspace,
                     (cudaStream_t)stream);
}

extern "C" int mcacq_log_hvi_backward(const double* grad_out, const double* out, const double* obj, const double* cell_lower,
                                      const double* cell_upper, int64_t B, int q, int m, int nc, double tau_relu, double tau_max,
                                      double* grad_obj, double* lcl_workspace, void* stream) {
  using namespace mcacq;
  if (!grad_out || !out) return MCACQ_EINVAL;
  int rc = hv_check(obj, cell_lower, cell_upper, grad_obj, lcl_workspace, B, q, m, nc, tau_relu, tau_max);
  if (rc) return rc;
  if (B == 0) return 0;
  g_launch_count = 0;
  return log_hvi_run(1, grad_out, out, obj, cell_lower, cell_upper, B, q, m, nc, tau_relu, tau_max, grad_obj, lcl_workspace,
                     (cudaStream_t)stream);
}
