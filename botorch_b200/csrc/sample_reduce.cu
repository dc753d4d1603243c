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

#include "sample_reduce_common.cuh"

namespace mcacq {

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
  // The rows a warp owns (warp, warp + NW, ...) advance TOGETHER through the substitution: every factor entry is loaded once
  // for all of them and their shuffle trees overlap, so the serial chain is r steps per warp instead of r per row.  Per row
  // the arithmetic (lane-strided partial sums, xor tree) is unchanged.
  {
    constexpr int RPW = (QMAX + NW - 1) / NW;
    for (int j = 0; j < r; j++) {
      const double* Lj = p.L_base + (size_t)j * r;
      double part[RPW];
#pragma unroll
      for (int t = 0; t < RPW; t++) part[t] = 0.0;
      for (int k = lane; k < j; k += 32) {
        const double lv = Lj[k];
#pragma unroll
        for (int t = 0; t < RPW; t++) {
          const int i = warp + t * NW;
          if (i < q) part[t] = fma(brow[(size_t)i * r + k], lv, part[t]);
        }
      }
#pragma unroll
      for (int t = 0; t < RPW; t++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part[t] += __shfl_xor_sync(0xffffffffu, part[t], o);
      if (lane == 0) {
        const double ljj = Lj[j];
#pragma unroll
        for (int t = 0; t < RPW; t++) {
          const int i = warp + t * NW;
          if (i < q) brow[(size_t)i * r + j] = (brow[(size_t)i * r + j] - part[t]) / ljj;
        }
      }
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
    else res = m + fm_log(s) - fm_log((double)S);
    if (s_info & MCACQ_INFO_NOT_PSD) res = CUDART_NAN;
    p.acq[bb] = res;
    p.info[bb] = s_info | (s_nonfinite ? MCACQ_INFO_NONFINITE : 0);
  }
}

// ---- host launchers --------------------------------------------------------------------------------
static size_t fwd_smem(int q, int r, int S, int wide) {
  int QP = (q + 1) & ~1;
  size_t scratch = (size_t)q * (r > q ? r : q);
  return ((size_t)(r + q) * QP + scratch + (size_t)q * q + q + 2 * SR_WARPS + q + (wide ? (size_t)S : 0)) * sizeof(double);
}

// dynamic shared memory of the (wide, i.e. larger) forward launch for this shape (host-side query: mcacq_fused_supported)
size_t sample_reduce_fwd_smem(int q, int r, int S) { return fwd_smem(q, r, S, 1); }

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


}  // namespace mcacq
