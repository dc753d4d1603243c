// Shared device helpers of the sample / reduce kernels (sample_reduce.cu: forward, sample_reduce_bwd.cu: backward; two
// translation units so that the template instantiations compile in parallel).
#pragma once
#include <cstdlib>
#include "common.cuh"
#include "params.cuh"
#include <math_constants.h>

namespace mcacq {



constexpr int SR_THREADS = 128;
constexpr int SR_WARPS = SR_THREADS / 32;

// ---- utility pieces -----------------------------------------------------------------------------
// Reciprocal to ~1 ulp without the IEEE division slow path: MUFU seed (>= 20 bits) + two Newton steps.
// Arguments here are always finite, normal and >= 1e-300 (1 + u^2, 2 + 2v + v^2, tau * f).
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

// log-improvement value and derivative w.r.t. z (the improvement y - best).
// Utility modes (mcacq_mc.fat):  0 log_softplus + smooth_amax + logmeanexp   (qLogEI / qLogNEI, fat=False)
//                                 1 log_fatplus  + fatmax      + logmeanexp   (qLogEI / qLogNEI, default)
//                                 2 relu(z)      + amax        + mean         (qEI / qNEI,  monte_carlo.py:427-437, 607-616)
//                                 3 y            + amax        + mean         (qSimpleRegret, :821-830; best = 0)
//                                 4 sigmoid(z/tau_relu) + amax + mean         (qProbabilityOfImprovement, :752-763)
template <bool GRAD>
__device__ __forceinline__ double log_improve(double z, double tau, double inv_tau, int fat, double& dli) {
  if (fat == 2) {
    if (GRAD) dli = (z > 0.0) ? 1.0 : 0.0;
    return fmax(z, 0.0);
  }
  if (fat == 3) {
    if (GRAD) dli = 1.0;
    return z;
  }
  if (fat == 4) {
    const double u = z * inv_tau;
    const double sg = (u >= 0.0) ? 1.0 / (1.0 + fm_exp(-u)) : fm_exp(u) / (1.0 + fm_exp(u));
    if (GRAD) dli = sg * (1.0 - sg) * inv_tau;
    return sg;
  }
  if (fat) {
    const double u = z * inv_tau;
    double sp, dsp;
    if (u > 20.0) {  // torch softplus threshold
      sp = u; dsp = 1.0;
    } else if (u < -746.0) {  // exp(u) == 0 exactly in fp64
      sp = 0.0; dsp = 0.0;
    } else {
      const double e = fm_exp(u);
      sp = fm_log1p_nonneg(e);
      dsp = e / (e + 1.0);
    }
    const double den = fma(u, u, 1.0);
    const double ca = (den < 1e300) ? fast_rcp(den) : 0.0;
    const double f = sp + 0.1 * ca;
    const double tf = tau * f;
    if (GRAD) dli = (dsp - 0.2 * u * ca * ca) * ((tf > 1e-300 && tf < 1e300) ? fast_rcp(tf) : 1.0 / tf);
    return fm_log(tf);
  } else {
    const double xt = z / tau;
    if (xt > -35.0) {
      const double beta = 1.0 / tau;
      const double xb = z * beta;
      double sp, dsp;
      if (xb > 32.0) { sp = z; dsp = 1.0; }
      else { const double e = fm_exp(xb); sp = fm_log1p_nonneg(e) / beta; dsp = e / (e + 1.0); }
      if (GRAD) dli = dsp / sp;
      return fm_log(sp);
    } else {
      if (GRAD) dli = 1.0 / tau;
      return xt + fm_log(tau);
    }
  }
}


// ---- smoothed feasibility of one sample value (botorch/utils/objective.py:135-211) -----------------------------------
// log sigmoid(u) / log fatmoid(u) and its derivative with respect to u.
__device__ __forceinline__ double log_feas_term(double u, int fat, double& dlf) {
  if (fat) {
    // fatmoid (safe_math.py:441-458): u < 0: 2/3 cauchy(u - 1/sqrt3), else 1 - 2/3 cauchy(u + 1/sqrt3)
    const double c3 = 0.57735026918962576451;
    if (u < 0.0) {
      const double a = u - c3, den = fma(a, a, 1.0);
      dlf = -2.0 * a / den;
      return fm_log((2.0 / 3.0) / den);
    }
    const double bb = u + c3, den = fma(bb, bb, 1.0);
    const double fv = 1.0 - (2.0 / 3.0) / den;
    dlf = ((4.0 / 3.0) * bb / (den * den)) / fv;
    return fm_log(fv);
  }
  // logexpit(u) = -log1pexp(-u) (safe_math.py:96-98, 78-93: log1p(exp(x)) for x <= 18, x + exp(-x) above)
  const double x = -u;
  double l1p, sig;   // log1pexp(x), sigmoid(x) = d log1pexp / dx
  if (x <= 18.0) { const double e = fm_exp(x); l1p = fm_log1p_nonneg(e); sig = e / (1.0 + e); }
  else { const double e = fm_exp(-x); l1p = x + e; sig = 1.0 - e; }
  dlf = sig;         // d(-log1pexp(-u)) / du = sigmoid(-u)
  return -l1p;
}

// Per-sample, per-point utility of a posterior sample value y: objective (affine), utility mode, constraint weighting.
// dy = d val / d y,  dm = d val / d mu_i (modes 5 / 6: mu_i = MC mean of the objective).
// PLAIN: the default qLogEI / qLogNEI configuration (utility mode 1, identity objective, no constraints) with the mode
// fields folded to constants, so that the other modes cost neither registers nor issue slots on the benchmark path.
template <bool GRAD, bool PLAIN>
__device__ __forceinline__ double sr_element(const SRParams& p, double yi, double bst, double mu, double inv_tau_relu,
                                             double& dy, double& dm) {
  if (PLAIN) {
    dm = 0.0;
    return log_improve<GRAD>(yi - bst, p.tau_relu, inv_tau_relu, 1, dy);
  }
  const double obj = fma(p.obj_w, yi, p.obj_o);
  double val, dobj = 0.0;
  dm = 0.0;
  if (p.fat >= 5) {
    const double dev = obj - mu;
    const double sgn = (dev > 0.0) ? 1.0 : ((dev < 0.0) ? -1.0 : 0.0);
    val = ((p.fat == 5) ? mu : 0.0) + p.util_param * fabs(dev);
    if (GRAD) { dobj = p.util_param * sgn; dm = ((p.fat == 5) ? 1.0 : 0.0) - p.util_param * sgn; }
  } else {
    val = log_improve<GRAD>(obj - bst, p.tau_relu, inv_tau_relu, p.fat, dobj);
  }
  dy = dobj * p.obj_w;
  if (p.n_con > 0) {
    double lf = 0.0, dlf = 0.0;
    for (int k = 0; k < p.n_con; k++) {
      double dk;
      lf += log_feas_term(-(fma(p.con_a[k], yi, p.con_b[k])) / p.con_eta[k], p.con_fat, dk);
      dlf += dk * (-p.con_a[k] / p.con_eta[k]);
    }
    if (p.fat <= 1) { val += lf; dy += dlf; }          // log family: add the log-indicator (monte_carlo.py:322-348)
    else { const double F = fm_exp(lf); dy = dy * F + val * F * dlf; dm *= F; val *= F; }
  }
  return val;
}

// q-reduction: fatmax (fat) or smooth_amax; optionally the weights d fm / d li_i.
template <int QMAX, bool GRAD>
__device__ __forceinline__ double q_reduce(const double (&li)[QMAX], int q, double tau, double inv_tau, int fat,
                                           double (&w)[QMAX]) {
  double M = -CUDART_INF;
#pragma unroll
  for (int i = 0; i < QMAX; i++) if (i < q) M = fmax(M, li[i]);
  if (fat >= 2) {  // torch.amax over q: gradient split evenly among ties
    if (GRAD) {
      int cnt = 0;
#pragma unroll
      for (int i = 0; i < QMAX; i++) if (i < q) cnt += (li[i] == M) ? 1 : 0;
      const double wgt = 1.0 / (double)(cnt > 0 ? cnt : 1);
#pragma unroll
      for (int i = 0; i < QMAX; i++) if (i < q) w[i] = (li[i] == M) ? wgt : 0.0;
    }
    return M;
  }
  if (isinf(M) || isnan(M)) {
    // _inf_max_helper: the result is the sum of the infinite maxima; gradient 1 on those entries
    double res = 0.0;
#pragma unroll
    for (int i = 0; i < QMAX; i++) if (i < q) {
      bool is_max = (li[i] == M);
      if (is_max) res += li[i];
      if (GRAD) w[i] = is_max ? 1.0 : 0.0;
    }
    return isnan(M) ? M : res;
  }
  if (fat) {
    double P = 0.0, dsum = 0.0;
    int cnt = 0;
    double dp[QMAX];
#pragma unroll
    for (int i = 0; i < QMAX; i++) if (i < q) {
      const double v = (M - li[i]) * inv_tau;
      const double den = fma(v, v + 2.0, 2.0);
      const double rden = (den < 1e300) ? fast_rcp(den) : 0.0;
      P += 2.0 * rden;
      if (GRAD) {
        dp[i] = -(4.0 + 4.0 * v) * rden * rden;
        dsum += dp[i];
        cnt += (li[i] == M) ? 1 : 0;
      }
    }
    if (GRAD) {
      const double head = (1.0 + dsum / P) / (double)cnt;
#pragma unroll
      for (int i = 0; i < QMAX; i++) if (i < q) w[i] = ((li[i] == M) ? head : 0.0) - dp[i] / P;
    }
    return M + tau * fm_log(P);
  } else {
    const double Mt = M / tau;
    double ssum = 0.0;
#pragma unroll
    for (int i = 0; i < QMAX; i++) if (i < q) {
      const double e = fm_exp(li[i] / tau - Mt);
      ssum += e;
      if (GRAD) w[i] = e;
    }
    if (GRAD) {
#pragma unroll
      for (int i = 0; i < QMAX; i++) if (i < q) w[i] /= ssum;
    }
    return (Mt + fm_log(ssum)) * tau;
  }
}

__device__ __forceinline__ void lse_push(double& m, double& s, double f) {
  if (f == -CUDART_INF) return;
  if (f > m) { s = s * fm_exp(m - f) + 1.0; m = f; }
  else s += fm_exp(f - m);
}
__device__ __forceinline__ void lse_merge(double& m, double& s, double m2, double s2) {
  const double M = fmax(m, m2);
  if (M == -CUDART_INF) { m = M; s = 0.0; return; }
  if (isinf(M)) { m = M; s = 1.0; return; }
  const double a = (m == -CUDART_INF) ? 0.0 : s * fm_exp(m - M);
  const double c = (m2 == -CUDART_INF) ? 0.0 : s2 * fm_exp(m2 - M);
  m = M; s = a + c;
}

// ---- shared prologue: load factors into shared memory -------------------------------------------
// coefT[j][i] (j < r: B[i][j]; j >= r: C[i][j-r]), row pitch QP = q rounded up to even.
__device__ __forceinline__ int coef_pitch(int q) { return (q + 1) & ~1; }

// q-batches below which the wide (4 x 128 threads) variants are used: fewer CTAs than the machine has SM sub-partitions
constexpr int64_t SR_WIDE_BELOW = 296;
static inline bool sr_use_wide(int64_t b) {
  const char* e = getenv("MCACQ_SR_WIDE");  // test hook: 0 / 1 force the narrow / wide variant (results must not differ)
  return e != nullptr ? atoi(e) != 0 : b < SR_WIDE_BELOW;
}

static inline bool sr_plain(const SRParams& p) {
  return p.fat == 1 && p.n_con == 0 && p.obj_w == 1.0 && p.obj_o == 0.0;
}


}  // namespace mcacq
