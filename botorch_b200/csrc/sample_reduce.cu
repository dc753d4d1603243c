// Fused reparameterised-sample / utility / reduction kernels (forward and hand-written backward).
//
// Per q-batch (one CTA):
//   B  = Sxb L_base^{-T}                         torch.linalg.solve_triangular in sample_cached_cholesky
//                                                (botorch/utils/low_rank.py:130-132)
//   C  = psd_safe_cholesky(Sxx - B B^T, 6)       (low_rank.py:137-140; jitter 1e-8*10^i on failure)
//   y[s][i] = mean[i] + sum_j B[i][j] Z[s][j] + sum_{j<=i} C[i][j] Z[s][r+j]      (low_rank.py:143-160)
//   li = log_fatplus(y - best[s], tau_relu)      (botorch/acquisition/logei.py:688-715, safe_math.py:298-325)
//   fm[s] = fatmax_i(li, tau_max)                (safe_math.py:328-355, _inf_max_helper :146-191)
//   acq = logsumexp_s fm[s] - log S              (safe_math.py:213-225)
// r == 0 gives the qLogEI path (posteriors/gpytorch.py:86-127: y = mean + chol(Sxx) z).
// The backward kernel recomputes the per-sample chain, forms the softmax/fatmax/log_fatplus weights
// (same weights as botorch/csrc/logei_fused.cpp:109-111, 143-174), contracts them with
// the base samples, and runs the q x q Cholesky reverse-mode and the triangular-solve reverse-mode in
// shared memory.
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
    const double sg = (u >= 0.0) ? 1.0 / (1.0 + exp(-u)) : exp(u) / (1.0 + exp(u));
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
      const double e = exp(u);
      sp = log1p(e);
      dsp = e / (e + 1.0);
    }
    const double den = fma(u, u, 1.0);
    const double ca = (den < 1e300) ? fast_rcp(den) : 0.0;
    const double f = sp + 0.1 * ca;
    const double tf = tau * f;
    if (GRAD) dli = (dsp - 0.2 * u * ca * ca) * ((tf > 1e-300 && tf < 1e300) ? fast_rcp(tf) : 1.0 / tf);
    return log(tf);
  } else {
    const double xt = z / tau;
    if (xt > -35.0) {
      const double beta = 1.0 / tau;
      const double xb = z * beta;
      double sp, dsp;
      if (xb > 32.0) { sp = z; dsp = 1.0; }
      else { const double e = exp(xb); sp = log1p(e) / beta; dsp = e / (e + 1.0); }
      if (GRAD) dli = dsp / sp;
      return log(sp);
    } else {
      if (GRAD) dli = 1.0 / tau;
      return xt + log(tau);
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
      return log((2.0 / 3.0) / den);
    }
    const double bb = u + c3, den = fma(bb, bb, 1.0);
    const double fv = 1.0 - (2.0 / 3.0) / den;
    dlf = ((4.0 / 3.0) * bb / (den * den)) / fv;
    return log(fv);
  }
  // logexpit(u) = -log1pexp(-u) (safe_math.py:96-98, 78-93: log1p(exp(x)) for x <= 18, x + exp(-x) above)
  const double x = -u;
  double l1p, sig;   // log1pexp(x), sigmoid(x) = d log1pexp / dx
  if (x <= 18.0) { const double e = exp(x); l1p = log1p(e); sig = e / (1.0 + e); }
  else { const double e = exp(-x); l1p = x + e; sig = 1.0 - e; }
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
    else { const double F = exp(lf); dy = dy * F + val * F * dlf; dm *= F; val *= F; }
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
    return M + tau * log(P);
  } else {
    const double Mt = M / tau;
    double ssum = 0.0;
#pragma unroll
    for (int i = 0; i < QMAX; i++) if (i < q) {
      const double e = exp(li[i] / tau - Mt);
      ssum += e;
      if (GRAD) w[i] = e;
    }
    if (GRAD) {
#pragma unroll
      for (int i = 0; i < QMAX; i++) if (i < q) w[i] /= ssum;
    }
    return (Mt + log(ssum)) * tau;
  }
}

__device__ __forceinline__ void lse_push(double& m, double& s, double f) {
  if (f == -CUDART_INF) return;
  if (f > m) { s = s * exp(m - f) + 1.0; m = f; }
  else s += exp(f - m);
}
__device__ __forceinline__ void lse_merge(double& m, double& s, double m2, double s2) {
  const double M = fmax(m, m2);
  if (M == -CUDART_INF) { m = M; s = 0.0; return; }
  if (isinf(M)) { m = M; s = 1.0; return; }
  const double a = (m == -CUDART_INF) ? 0.0 : s * exp(m - M);
  const double c = (m2 == -CUDART_INF) ? 0.0 : s2 * exp(m2 - M);
  m = M; s = a + c;
}

// ---- shared prologue: load factors into shared memory -------------------------------------------
// coefT[j][i] (j < r: B[i][j]; j >= r: C[i][j-r]), row pitch QP = q rounded up to even.
__device__ __forceinline__ int coef_pitch(int q) { return (q + 1) & ~1; }

// Forward -------------------------------------------------------------------------------------------
// W > 1 ("wide", for the few q-batches of an optimiser round): W x 128 threads evaluate the per-sample utilities, park them
// in shared memory, and the first 128 threads then fold them in exactly the order of the W = 1 kernel -- results are
// bit-identical for every W, only the latency per q-batch changes.
template <int QMAX, int NS, int W, bool PLAIN>
__global__ void __launch_bounds__(SR_THREADS * W)
sample_reduce_fwd_kernel(SRParams p) {
  constexpr int NT = SR_THREADS * W, NW = SR_WARPS * W;
  const int fat = PLAIN ? 1 : p.fat;
  extern __shared__ __align__(16) double sm[];
  const int q = p.q, r = p.r, S = p.S;
  const int QP = coef_pitch(q);
  double* coefT = sm;                          // [(r+q)][QP]
  double* brow = coefT + (size_t)(r + q) * QP; // [q][r]   row-major scratch for the triangular solve
  double* Tm = brow + (size_t)q * (r > q ? r : q);  // [q][q]  (scratch is q*max(r,q): it later holds the q x q factor,
                                                    //  which must not alias T or the jitter retries would re-read garbage)
  double* smean = Tm + q * q;                  // [q]
  double* red = smean + q;                     // [2 * SR_WARPS]
  double* smu = red + 2 * SR_WARPS;            // [q] MC mean of the objective (utility modes 5 / 6)
  double* fmv = smu + q;                       // [S] per-sample utilities (W > 1 only)
  __shared__ int s_info;
  __shared__ int s_nonfinite;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t bb = blockIdx.x;

  for (int idx = tid; idx < q * r; idx += NT) brow[idx] = p.Sxb[bb * q * r + idx];
  for (int idx = tid; idx < q; idx += NT) smean[idx] = p.mean[bb * q + idx];
  for (int idx = tid; idx < (r + q) * QP; idx += NT) coefT[idx] = 0.0;
  if (tid == 0) { s_info = 0; s_nonfinite = 0; }
  __syncthreads();

  // ---- B = Sxb L^{-T}: forward substitution, one warp per row, lanes over the dot product
  for (int i = warp; i < q; i += NW) {
    double* bi = brow + (size_t)i * r;
    for (int j = 0; j < r; j++) {
      const double* Lj = p.L_base + (size_t)j * r;
      double part = 0.0;
      for (int k = lane; k < j; k += 32) part = fma(bi[k], Lj[k], part);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if (lane == 0) bi[j] = (bi[j] - part) / Lj[j];
      __syncwarp();
    }
  }
  __syncthreads();
  for (int idx = tid; idx < q * r; idx += NT) {
    const int i = idx / r, j = idx - i * r;
    const double v = brow[idx];
    coefT[j * QP + i] = v;
    p.Bm[bb * q * r + idx] = v;
  }
  // ---- T = Sxx - B B^T (lower triangle)
  for (int idx = tid; idx < q * q; idx += NT) {
    const int i = idx / q, j = idx - i * q;
    double v = 0.0;
    if (j <= i) {
      double dot = 0.0;
      for (int k = 0; k < r; k++) dot = fma(brow[i * r + k], brow[j * r + k], dot);
      v = p.Sxx[bb * q * q + idx] - dot;
    }
    Tm[idx] = v;
  }
  __syncthreads();

  // ---- C = psd_safe_cholesky(T): warp 0, lane = row; up to 6 jitter escalations
  if (warp == 0) {
    double* Cs = brow;  // reuse scratch (q*q <= needs q*max(r,q); sized on host)
    double jit_prev = 0.0;
    double diag = (lane < q) ? Tm[lane * q + lane] : 1.0;
    int info = 0;
    bool ok = false;
    for (int attempt = 0; attempt <= 6; attempt++) {
      if (attempt > 0) {
        // linear_operator psd_safe_cholesky: jitter_new = 1e-8 * (10 ** i), i = attempt - 1
        const double p10[6] = {1.0, 10.0, 100.0, 1000.0, 10000.0, 100000.0};
        const double jit_new = 1e-8 * p10[attempt - 1];
        diag += p.jitter_f32 ? (double)(float)(jit_new - jit_prev) : (jit_new - jit_prev);
        jit_prev = jit_new;
        info = attempt;
      }
      ok = true;
      for (int j = 0; j < q; j++) {
        double v = 0.0;
        if (lane >= j && lane < q) {
          v = (lane == j) ? diag : Tm[lane * q + j];
          for (int k = 0; k < j; k++) v -= Cs[lane * q + k] * Cs[j * q + k];
        }
        const double dj = __shfl_sync(0xffffffffu, v, j);
        if (!(dj > 0.0)) { ok = false; __syncwarp(); break; }  // (orders this attempt's reads before the retry's writes)
        const double sj = sqrt(dj);
        if (lane >= j && lane < q) Cs[lane * q + j] = (lane == j) ? sj : v / sj;
        __syncwarp();
      }
      if (ok) break;
    }
    if (!ok) info = 6 | MCACQ_INFO_NOT_PSD;
    {
      // conditioning of the q-batch: rho = min_i C_ii^2 / Sxx_ii, the smallest fraction of a point's posterior variance that
      // is left after conditioning on the baseline draws and on the preceding points of the q-batch (1 = independent points,
      // -> 0 = the joint covariance is singular); floor(-4 log2 rho), saturated at 255, goes into bits 8..15 of the status
      // word.  Rounding errors of the posterior blocks reach the value and its gradient amplified by 1 / rho, which is what
      // the Python layer uses to route ill-conditioned q-batches of the int8 contraction mode to the FP64 contraction.
      double rho = 1.0;
      if (lane < q) {
        const double sii = p.Sxx[bb * q * q + lane * q + lane];
        const double cii = ok ? Cs[lane * q + lane] : 0.0;
        rho = (sii > 0.0) ? cii * cii / sii : 0.0;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) rho = fmin(rho, __shfl_xor_sync(0xffffffffu, rho, o));
      int cond = 255;
      if (rho >= 1.0) cond = 0;
      else if (rho > 0.0) { const double c = -4.0 * log2(rho); cond = c < 255.0 ? (int)c : 255; }
      info |= cond << MCACQ_INFO_COND_SHIFT;
      // variance collapse of the q-batch: floor(-4 log2 min_i Sxx_ii / prior), saturated at 255, in bits 16..23.  Sxx = prior
      // - |a|^2 is a difference: a contraction that is accurate to a FRACTION OF THE PRIOR (the int8 mode) reproduces Sxx_ii
      // to that fraction divided by Sxx_ii / prior, which is what the Python layer bounds per fitted model.
      int vbyte = 0;
      if (p.prior_var > 0.0) {
        double vr = 1.0;
        if (lane < q) {
          const double sii = p.Sxx[bb * q * q + lane * q + lane];
          vr = (sii > 0.0) ? sii / p.prior_var : 0.0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) vr = fmin(vr, __shfl_xor_sync(0xffffffffu, vr, o));
        vbyte = 255;
        if (vr >= 1.0) vbyte = 0;
        else if (vr > 0.0) { const double c = -4.0 * log2(vr); vbyte = c < 255.0 ? (int)c : 255; }
      }
      info |= vbyte << MCACQ_INFO_VAR_SHIFT;
    }
    if (lane < q) {
      for (int j = 0; j < q; j++) {
        double v = (j <= lane) ? (ok ? Cs[lane * q + j] : CUDART_NAN) : 0.0;
        coefT[(r + j) * QP + lane] = v;
        p.Cm[bb * q * q + lane * q + j] = v;
      }
    }
    if (lane == 0) s_info = info;
  }
  __syncthreads();

  // MC mean of the objective per point in closed form: obj_w (mean_i + sum_j coef_ij Zbar_j) + obj_o
  if (fat >= 5) {
    for (int i = tid; i < q; i += NT) {
      double m = smean[i];
      for (int j = 0; j < r + q; j++) m = fma(coefT[j * QP + i], p.Zbar[j], m);
      smu[i] = fma(p.obj_w, m, p.obj_o);
    }
  }
  __syncthreads();
  // ---- samples: NS samples per thread, coefficients broadcast from shared memory
  const double inv_tau_relu = 1.0 / p.tau_relu, inv_tau_max = 1.0 / p.tau_max;
  double lm = -CUDART_INF, ls = 0.0;
  bool nonfinite = false;
  if (W > 1) {
    static_assert(W == 1 || NS == 1, "the wide kernel evaluates one sample per thread and pass");
    for (int s0 = tid; s0 < S; s0 += NT) {
      double y[QMAX];
#pragma unroll
      for (int i = 0; i < QMAX; i++) y[i] = 0.0;
#pragma unroll 4
      for (int j = 0; j < r + q; j++) {
        const double z = p.Zt[(size_t)j * S + s0];
        const double* cj = coefT + j * QP;
#pragma unroll
        for (int i = 0; i < QMAX; i += 2) {
          if (i < q) {
            const double2 c2 = *reinterpret_cast<const double2*>(cj + i);
            y[i] = fma(c2.x, z, y[i]);
            if (i + 1 < QMAX) y[i + 1] = fma(c2.y, z, y[i + 1]);
          }
        }
      }
      const double bst = p.best[s0];
      double li[QMAX], wdummy[QMAX];
#pragma unroll
      for (int i = 0; i < QMAX; i++) {
        if (i < q) {
          const double yi = y[i] + smean[i];
          if (!isfinite(yi)) nonfinite = true;
          double dl, dmm;
          li[i] = sr_element<false, PLAIN>(p, yi, bst, (fat >= 5) ? smu[i] : 0.0, inv_tau_relu, dl, dmm);
        } else li[i] = -CUDART_INF;
      }
      fmv[s0] = q_reduce<QMAX, false>(li, q, p.tau_max, inv_tau_max, fat, wdummy);
    }
    if (nonfinite) s_nonfinite = 1;
    __syncthreads();
    if (tid < SR_THREADS) {
      for (int s0 = tid; s0 < S; s0 += SR_THREADS) {   // the W = 1 kernel's per-thread sample order
        const double fm = fmv[s0];
        if (fat >= 2) ls += fm;
        else if (fm == CUDART_INF) { lm = fm; ls = 1.0; }
        else lse_push(lm, ls, fm);
      }
    }
  } else {
  for (int s0 = tid * NS; s0 < S; s0 += SR_THREADS * NS) {
      double y[NS][QMAX];
#pragma unroll
      for (int ns = 0; ns < NS; ns++)
#pragma unroll
        for (int i = 0; i < QMAX; i++) y[ns][i] = 0.0;
#pragma unroll 4
      for (int j = 0; j < r + q; j++) {
        double z[NS];
#pragma unroll
        for (int ns = 0; ns < NS; ns++) z[ns] = (s0 + ns < S) ? p.Zt[(size_t)j * S + s0 + ns] : 0.0;
        const double* cj = coefT + j * QP;
#pragma unroll
        for (int i = 0; i < QMAX; i += 2) {
          if (i < q) {
            const double2 c2 = *reinterpret_cast<const double2*>(cj + i);
#pragma unroll
            for (int ns = 0; ns < NS; ns++) {
              y[ns][i] = fma(c2.x, z[ns], y[ns][i]);
              if (i + 1 < QMAX) y[ns][i + 1] = fma(c2.y, z[ns], y[ns][i + 1]);
            }
          }
        }
      }
#pragma unroll
      for (int ns = 0; ns < NS; ns++) {
        if (s0 + ns < S) {
          const double bst = p.best[s0 + ns];
          double li[QMAX], wdummy[QMAX];
#pragma unroll
          for (int i = 0; i < QMAX; i++) {
            if (i < q) {
              const double yi = y[ns][i] + smean[i];
              if (!isfinite(yi)) nonfinite = true;
              double dl, dmm;
              li[i] = sr_element<false, PLAIN>(p, yi, bst, (fat >= 5) ? smu[i] : 0.0, inv_tau_relu, dl, dmm);
            } else li[i] = -CUDART_INF;
          }
          const double fm = q_reduce<QMAX, false>(li, q, p.tau_max, inv_tau_max, fat, wdummy);
          if (fat >= 2) ls += fm;  // plain mean over the samples
          else if (fm == CUDART_INF) { lm = fm; ls = 1.0; }
          else lse_push(lm, ls, fm);
        }
      }
    }
  }
  if (nonfinite) s_nonfinite = 1;
  // ---- CTA logsumexp (modes 0/1) or sum (modes >= 2), fixed combination order
  if (warp < SR_WARPS) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double m2 = __shfl_xor_sync(0xffffffffu, lm, o);
      const double s2 = __shfl_xor_sync(0xffffffffu, ls, o);
      if (fat >= 2) ls += s2;
      else lse_merge(lm, ls, m2, s2);
    }
    if (lane == 0) { red[2 * warp] = lm; red[2 * warp + 1] = ls; }
  }
  __syncthreads();
  if (tid == 0) {
    double m = red[0], s = red[1];
    for (int w = 1; w < SR_WARPS; w++) {
      if (fat >= 2) s += red[2 * w + 1];
      else lse_merge(m, s, red[2 * w], red[2 * w + 1]);
    }
    double res;
    if (fat >= 2) res = s / (double)S;
    else if (S == 0 || m == -CUDART_INF) res = -CUDART_INF;
    else if (isinf(m)) res = m;
    else res = m + log(s) - log((double)S);
    if (s_info & MCACQ_INFO_NOT_PSD) res = CUDART_NAN;
    p.acq[bb] = res;
    p.info[bb] = s_info | (s_nonfinite ? MCACQ_INFO_NONFINITE : 0);
  }
}

// Backward ------------------------------------------------------------------------------------------
// W > 1: the per-sample weights (pass A, the expensive part) are evaluated by W x 128 threads; the contraction with the
// base samples (pass B) keeps the 4-warp split of the W = 1 kernel, so the results are bit-identical for every W.
template <int QMAX, int NS, int W, bool PLAIN>
__global__ void __launch_bounds__(SR_THREADS * W)
sample_reduce_bwd_kernel(SRParams p, int chunk) {
  constexpr int NT = SR_THREADS * W, NW = SR_WARPS * W;
  const int fat = PLAIN ? 1 : p.fat;
  extern __shared__ __align__(16) double sm[];
  const int q = p.q, r = p.r, S = p.S;
  const int QP = coef_pitch(q);
  const int NC = r + q + 1;                    // contraction outputs per row: B (r), C (q), mean (1)
  double* coefT = sm;                          // [(r+q)][QP]
  double* gco = coefT + (size_t)(r + q) * QP;  // [q][NC]  accumulated d/d[B C mean]
  double* smean = gco + (size_t)q * NC;        // [q]
  double* mats = smean + q;                    // 4 x [q][q] scratch: L, P/X, G, gT
  const int GP = q | 1;                        // odd pitch for the weight rows
  double* gy = mats + 4 * q * q;               // [chunk + 3][GP]  per-sample weights (zero padded to a multiple of 4 rows)
  double* stage = gy + (size_t)(chunk + 4) * GP;  // [SR_WARPS][QMAX][8] cross-warp staging of DMMA partials
  double* smu = stage + (size_t)SR_WARPS * QMAX * 8;   // [q] MC mean of the objective (modes 5 / 6)
  double* sam = smu + q;                               // [q] sum_s d acq / d mu_i
  double* gy2 = sam + q;                               // [chunk + 4][GP] per-sample d acq / d mu_i (modes 5 / 6 only)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t bb = blockIdx.x;

  for (int idx = tid; idx < (r + q) * QP; idx += NT) coefT[idx] = 0.0;
  for (int idx = tid; idx < q * NC; idx += NT) gco[idx] = 0.0;
  for (int idx = tid; idx < q; idx += NT) smean[idx] = p.mean[bb * q + idx];
  __syncthreads();
  for (int idx = tid; idx < q * r; idx += NT) {
    const int i = idx / r, j = idx - i * r;
    coefT[j * QP + i] = p.Bm[bb * q * r + idx];
  }
  for (int idx = tid; idx < q * q; idx += NT) {
    const int i = idx / q, j = idx - i * q;
    coefT[(r + j) * QP + i] = p.Cm[bb * q * q + idx];
  }
  __syncthreads();

  const double gout = p.grad_acq[bb];
  const double lse_total = p.acq[bb] + log((double)S);
  const double inv_tau_relu = 1.0 / p.tau_relu, inv_tau_max = 1.0 / p.tau_max;
  const bool mc_mean = fat >= 5;
  if (mc_mean) {
    for (int i = tid; i < q; i += NT) {
      double m = smean[i];
      for (int j = 0; j < r + q; j++) m = fma(coefT[j * QP + i], p.Zbar[j], m);
      smu[i] = fma(p.obj_w, m, p.obj_o);
      sam[i] = 0.0;
    }
    __syncthreads();
  }



  for (int c0 = 0; c0 < S; c0 += chunk) {
    const int cend = (c0 + chunk < S) ? c0 + chunk : S;
    // ---- pass A: per-sample weights gy[s][i]
    for (int s0 = c0 + tid * NS; s0 < cend; s0 += NT * NS) {
      double y[NS][QMAX];
#pragma unroll
      for (int ns = 0; ns < NS; ns++)
#pragma unroll
        for (int i = 0; i < QMAX; i++) y[ns][i] = 0.0;
#pragma unroll 4
      for (int j = 0; j < r + q; j++) {
        double z[NS];
#pragma unroll
        for (int ns = 0; ns < NS; ns++) z[ns] = (s0 + ns < cend) ? p.Zt[(size_t)j * S + s0 + ns] : 0.0;
        const double* cj = coefT + j * QP;
#pragma unroll
        for (int i = 0; i < QMAX; i += 2) {
          if (i < q) {
            const double2 c2 = *reinterpret_cast<const double2*>(cj + i);
#pragma unroll
            for (int ns = 0; ns < NS; ns++) {
              y[ns][i] = fma(c2.x, z[ns], y[ns][i]);
              if (i + 1 < QMAX) y[ns][i + 1] = fma(c2.y, z[ns], y[ns][i + 1]);
            }
          }
        }
      }
#pragma unroll
      for (int ns = 0; ns < NS; ns++) {
        if (s0 + ns < cend) {
          const double bst = p.best[s0 + ns];
          double li[QMAX], dli[QMAX], dmu[QMAX], w[QMAX];
#pragma unroll
          for (int i = 0; i < QMAX; i++) {
            if (i < q) li[i] = sr_element<true, PLAIN>(p, y[ns][i] + smean[i], bst, mc_mean ? smu[i] : 0.0, inv_tau_relu, dli[i], dmu[i]);
            else { li[i] = -CUDART_INF; dli[i] = 0.0; dmu[i] = 0.0; }
          }
          const double fm = q_reduce<QMAX, true>(li, q, p.tau_max, inv_tau_max, fat, w);
          double ws;
          if (fat >= 2) ws = gout / (double)S;
          else if (isinf(fm)) ws = (fm > 0) ? gout : ((isinf(lse_total) && lse_total < 0) ? gout : 0.0);
          else ws = gout * exp(fm - lse_total);
          double* gys = gy + (size_t)(s0 + ns - c0) * GP;
#pragma unroll
          for (int i = 0; i < QMAX; i++) if (i < q) gys[i] = ws * w[i] * dli[i];
          if (mc_mean) {
            double* gms = gy2 + (size_t)(s0 + ns - c0) * GP;
#pragma unroll
            for (int i = 0; i < QMAX; i++) if (i < q) gms[i] = ws * w[i] * dmu[i];
          }
        }
      }
    }
    __syncthreads();
    // zero the padding rows so that the k-loop can run in steps of 4
    {
      const int nloc = cend - c0;
      const int npad = (nloc + 3) & ~3;
      for (int idx = tid; idx < (npad - nloc) * GP; idx += NT) gy[(size_t)nloc * GP + idx] = 0.0;
    }
    __syncthreads();
    // ---- pass B (tensor pipe): gco[i][j] += sum_s gy[s][i] * Z[s][j]   (j == r+q: the mean column, Z == 1)
    //      DMMA.8x8x4 with A = gy^T (m = i, k = sample), B = Z (k = sample, n = j); the 4 warps split the samples,
    //      partial tiles are combined in a fixed order (deterministic).
    {
      const int g = lane >> 2, t4 = lane & 3;
      const int nloc = cend - c0;
      const int ksteps = (nloc + 3) >> 2;
      const int per_warp = (ksteps + SR_WARPS - 1) / SR_WARPS;
      const int k_begin = warp * per_warp;
      const int k_end = (k_begin + per_warp < ksteps) ? k_begin + per_warp : ksteps;
      const int ntj = (NC + 7) >> 3;
      constexpr int MT = QMAX / 8;
      for (int nj = 0; nj < ntj; nj++) {
        const int j = nj * 8 + g;
        const double* zr = (j < r + q) ? p.Zt + (size_t)j * S + c0 : nullptr;
        const double bconst = (j == r + q) ? 1.0 : 0.0;
        if (warp < SR_WARPS) {
          double acc[MT][2];
#pragma unroll
          for (int mi = 0; mi < MT; mi++) { acc[mi][0] = 0.0; acc[mi][1] = 0.0; }
          // four k-steps per trip with all operand loads issued first (the base samples come straight from L2)
          for (int kk = k_begin; kk < k_end; kk += 4) {
            double bv[4], av[4][MT];
#pragma unroll
            for (int u = 0; u < 4; u++) {
              const int sl = 4 * (kk + u) + t4;
              const bool live = (kk + u < k_end) && sl < nloc;
              bv[u] = (zr != nullptr) ? (live ? zr[sl] : 0.0) : ((kk + u < k_end) ? bconst : 0.0);
#pragma unroll
              for (int mi = 0; mi < MT; mi++) {
                const int i = mi * 8 + g;
                av[u][mi] = (live && i < q) ? gy[(size_t)sl * GP + i] : 0.0;
              }
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
              for (int mi = 0; mi < MT; mi++)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(acc[mi][0]), "+d"(acc[mi][1]) : "d"(av[u][mi]), "d"(bv[u]));
          }
#pragma unroll
          for (int mi = 0; mi < MT; mi++) {
            stage[((warp * QMAX) + mi * 8 + g) * 8 + 2 * t4] = acc[mi][0];
            stage[((warp * QMAX) + mi * 8 + g) * 8 + 2 * t4 + 1] = acc[mi][1];
          }
        }
        __syncthreads();
        for (int o = tid; o < QMAX * 8; o += NT) {
          const int i = o >> 3, jj = nj * 8 + (o & 7);
          if (i < q && jj < NC) {
            double a = 0.0;
#pragma unroll
            for (int w = 0; w < SR_WARPS; w++) a += stage[(w * QMAX + i) * 8 + (o & 7)];
            gco[i * NC + jj] += a;
          }
        }
        __syncthreads();
      }
      if (mc_mean) {
        // sum over the chunk's samples of d acq / d mu_i: the first 4 warps take the points, lanes stride over the samples
        // (the same split for every W, so wide and narrow launches agree bit for bit)
        if (warp < SR_WARPS) {
          for (int i = warp; i < q; i += SR_WARPS) {
            double a = 0.0;
            for (int sl = lane; sl < nloc; sl += 32) a += gy2[(size_t)sl * GP + i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0) sam[i] += a;
          }
        }
        __syncthreads();
      }
    }
  }
  __syncthreads();
  if (mc_mean) {
    // mu_i = obj_w (mean_i + sum_j coef_ij Zbar_j) + obj_o feeds back into the mean and into every coefficient
    for (int idx = tid; idx < q * NC; idx += NT) {
      const int i = idx / NC, j = idx - i * NC;
      gco[idx] += p.obj_w * sam[i] * ((j < r + q) ? p.Zbar[j] : 1.0);
    }
    __syncthreads();
  }

  // ---- Cholesky reverse-mode (warp 0): gT = sym( L^{-T} Phi(L^T gL) L^{-1} )
  double* Lm = mats;
  double* Xm = mats + q * q;
  double* Gm = mats + 2 * q * q;
  double* gT = mats + 3 * q * q;
  for (int idx = tid; idx < q * q; idx += NT) {
    const int i = idx / q, j = idx - i * q;
    Lm[idx] = (j <= i) ? coefT[(r + j) * QP + i] : 0.0;
  }
  __syncthreads();
  if (warp == 0) {
    // P = Phi(L^T gL), lane = column c
    if (lane < q) {
      const int c = lane;
      for (int a = 0; a < q; a++) {
        double v = 0.0;
        if (a >= c) {
          for (int k = a; k < q; k++) v = fma(Lm[k * q + a], gco[k * NC + r + c], v);  // gL[k][c], k >= a >= c
          if (a == c) v *= 0.5;
        }
        Xm[a * q + c] = v;
      }
      // X = L^{-T} P (back substitution down the rows), column c independent
      for (int a = q - 1; a >= 0; a--) {
        double v = Xm[a * q + c];
        for (int k = a + 1; k < q; k++) v -= Lm[k * q + a] * Xm[k * q + c];
        Xm[a * q + c] = v / Lm[a * q + a];
      }
    }
    __syncwarp();
    // G = X L^{-1}: lane = row a
    if (lane < q) {
      const int a = lane;
      for (int c = q - 1; c >= 0; c--) {
        double v = Xm[a * q + c];
        for (int k = c + 1; k < q; k++) v -= Gm[a * q + k] * Lm[k * q + c];
        Gm[a * q + c] = v / Lm[c * q + c];
      }
    }
    __syncwarp();
    for (int idx = lane; idx < q * q; idx += 32) {
      const int i = idx / q, j = idx - i * q;
      gT[idx] = 0.5 * (Gm[i * q + j] + Gm[j * q + i]);
    }
  }
  __syncthreads();
  // ---- outputs: gmean, gSxx = gT, gB_tot = gB - 2 gT B
  for (int idx = tid; idx < q; idx += NT) p.gmean[bb * q + idx] = gco[idx * NC + r + q];
  for (int idx = tid; idx < q * q; idx += NT) p.gSxx[bb * q * q + idx] = gT[idx];
  for (int idx = tid; idx < q * r; idx += NT) {
    const int i = idx / r, j = idx - i * r;
    double v = gco[i * NC + j];
    for (int k = 0; k < q; k++) v -= 2.0 * gT[i * q + k] * coefT[j * QP + k];
    gy[idx] = v;  // reuse gy scratch as gB_tot [q][r]  (chunk*GP >= q*r guaranteed by the host)
  }
  __syncthreads();
  // ---- gSxb = gB_tot L_base^{-1}: backward substitution, one warp per row
  for (int i = warp; i < q; i += NW) {
    double* gi = gy + (size_t)i * r;
    for (int j = r - 1; j >= 0; j--) {
      double part = 0.0;
      for (int k = j + 1 + lane; k < r; k += 32) part = fma(gi[k], p.L_base[(size_t)k * r + j], part);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if (lane == 0) gi[j] = (gi[j] - part) / p.L_base[(size_t)j * r + j];
      __syncwarp();
    }
    for (int j = lane; j < r; j += 32) p.gSxb[(bb * q + i) * r + j] = gi[j];
  }
}

// ---- host launchers --------------------------------------------------------------------------------
static size_t fwd_smem(int q, int r, int S, int wide) {
  int QP = (q + 1) & ~1;
  size_t scratch = (size_t)q * (r > q ? r : q);
  return ((size_t)(r + q) * QP + scratch + (size_t)q * q + q + 2 * SR_WARPS + q + (wide ? (size_t)S : 0)) * sizeof(double);
}

// q-batches below which the wide (4 x 128 threads) variants are used: fewer CTAs than the machine has SM sub-partitions
constexpr int64_t SR_WIDE_BELOW = 296;
static bool sr_use_wide(int64_t b) {
  const char* e = getenv("MCACQ_SR_WIDE");  // test hook: 0 / 1 force the narrow / wide variant (results must not differ)
  return e != nullptr ? atoi(e) != 0 : b < SR_WIDE_BELOW;
}

static size_t bwd_smem(int q, int r, int chunk, int qmax, bool mc_mean) {
  int QP = (q + 1) & ~1;
  int GP = q | 1;
  return ((size_t)(r + q) * QP + (size_t)q * (r + q + 1) + q + 4 * (size_t)q * q + (size_t)(chunk + 4) * GP +
          (size_t)SR_WARPS * qmax * 8 + 2 * (size_t)q + (mc_mean ? (size_t)(chunk + 4) * GP : 0)) * sizeof(double);
}

static inline bool sr_plain(const SRParams& p) {
  return p.fat == 1 && p.n_con == 0 && p.obj_w == 1.0 && p.obj_o == 0.0;
}

template <int QMAX, int NS, int W, bool PLAIN = false>
static int launch_sr_fwd(const SRParams& p, cudaStream_t st) {
  size_t smem = fwd_smem(p.q, p.r, p.S, W > 1);
  if (smem > 200 * 1024) return MCACQ_ELIMIT;
  auto kern = sample_reduce_fwd_kernel<QMAX, NS, W, PLAIN>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<<<(unsigned)p.b, SR_THREADS * W, smem, st>>>(p);
  count_launch();
  MCACQ_CUDA_CHECK_LAUNCH();
  return 0;
}

template <int QMAX, int NS, int W, bool PLAIN = false>
static int launch_sr_bwd(const SRParams& p, cudaStream_t st) {
  // chunk of samples whose weights are staged in shared memory (<= ~64 KB), at least r rows for the solve scratch
  const bool mc_mean = p.fat >= 5;
  int chunk = ((mc_mean ? 18 : 36) * 1024) / (8 * (p.q | 1));
  if (chunk > p.S) chunk = p.S;
  chunk = (chunk / (SR_THREADS * NS)) * (SR_THREADS * NS);
  if (chunk < SR_THREADS * NS) chunk = SR_THREADS * NS;
  if ((int64_t)chunk * (p.q | 1) < (int64_t)p.q * p.r) chunk = (p.q * p.r + (p.q | 1) - 1) / (p.q | 1);
  size_t smem = bwd_smem(p.q, p.r, chunk, QMAX, mc_mean);
  if (smem > 200 * 1024) return MCACQ_ELIMIT;
  auto kern = sample_reduce_bwd_kernel<QMAX, NS, W, PLAIN>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<<<(unsigned)p.b, SR_THREADS * W, smem, st>>>(p, chunk);
  count_launch();
  MCACQ_CUDA_CHECK_LAUNCH();
  return 0;
}

// out[0] = OR of the flag bits, out[1] = largest conditioning byte, out[2] = largest variance-collapse byte of info[0..b)
__global__ void info_summary_kernel(const int32_t* __restrict__ info, int64_t b, int32_t* __restrict__ out) {
  int fl = 0, c = 0, v = 0;
  for (int64_t i = threadIdx.x; i < b; i += blockDim.x) {
    const int w = info[i];
    fl |= w & MCACQ_INFO_FLAG_MASK;
    c = max(c, (w & MCACQ_INFO_COND_MASK) >> MCACQ_INFO_COND_SHIFT);
    v = max(v, (w & MCACQ_INFO_VAR_MASK) >> MCACQ_INFO_VAR_SHIFT);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    fl |= __shfl_xor_sync(0xffffffffu, fl, o);
    c = max(c, __shfl_xor_sync(0xffffffffu, c, o));
    v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  }
  __shared__ int sf[32], sc[32], sv[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sf[warp] = fl; sc[warp] = c; sv[warp] = v; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) { fl |= sf[w]; c = max(c, sc[w]); v = max(v, sv[w]); }
    out[0] = fl; out[1] = c; out[2] = v;
  }
}

int info_summary(const int32_t* info, int64_t b, int32_t* out, cudaStream_t st) {
  info_summary_kernel<<<1, 1024, 0, st>>>(info, b, out);
  count_launch();
  MCACQ_CUDA_CHECK_LAUNCH();
  return 0;
}

int sample_reduce_fwd(const SRParams& p, cudaStream_t st) {
  if (p.q <= 0 || p.q > MCACQ_MAX_Q || p.r < 0 || p.S <= 0) return MCACQ_ELIMIT;
  if (p.b == 0) return 0;
  // one sample per thread and pass: more resident warps beat per-thread ILP here (the transcendental chains do not
  // interleave across samples): 1.55 -> 0.96 ms forward, 2.86 -> 1.83 ms backward per C3 chunk
  const bool wide = sr_use_wide(p.b);
  if (p.q <= 8 && sr_plain(p)) return wide ? launch_sr_fwd<8, 1, 4, true>(p, st) : launch_sr_fwd<8, 1, 1, true>(p, st);
  if (p.q <= 8) return wide ? launch_sr_fwd<8, 1, 4>(p, st) : launch_sr_fwd<8, 1, 1>(p, st);
  if (p.q <= 16) return wide ? launch_sr_fwd<16, 1, 4>(p, st) : launch_sr_fwd<16, 1, 1>(p, st);
  return launch_sr_fwd<32, 1, 1>(p, st);
}

int sample_reduce_bwd(const SRParams& p, cudaStream_t st) {
  if (p.q <= 0 || p.q > MCACQ_MAX_Q || p.r < 0 || p.S <= 0) return MCACQ_ELIMIT;
  if (p.b == 0) return 0;
  const bool wide = sr_use_wide(p.b);
  if (p.q <= 8 && sr_plain(p)) return wide ? launch_sr_bwd<8, 1, 4, true>(p, st) : launch_sr_bwd<8, 1, 1, true>(p, st);
  if (p.q <= 8) return wide ? launch_sr_bwd<8, 1, 4>(p, st) : launch_sr_bwd<8, 1, 1>(p, st);
  if (p.q <= 16) return launch_sr_bwd<16, 1, 1>(p, st);
  return launch_sr_bwd<32, 1, 1>(p, st);
}

}  // namespace mcacq
