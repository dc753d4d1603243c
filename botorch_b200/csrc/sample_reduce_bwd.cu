// Backward of the fused sample / reduce stage (see sample_reduce.cu for the forward and the data layout).
#include "sample_reduce_common.cuh"

namespace mcacq {

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
          else ws = gout * fm_exp(fm - lse_total);
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
  //      (the rows of a warp advance together: each -- strided -- factor entry is loaded once for all of them; per row the
  //      arithmetic is unchanged, see the forward kernel)
  {
    constexpr int RPW = (QMAX + NW - 1) / NW;
    for (int j = r - 1; j >= 0; j--) {
      double part[RPW];
#pragma unroll
      for (int t = 0; t < RPW; t++) part[t] = 0.0;
      for (int k = j + 1 + lane; k < r; k += 32) {
        const double lv = p.L_base[(size_t)k * r + j];
#pragma unroll
        for (int t = 0; t < RPW; t++) {
          const int i = warp + t * NW;
          if (i < q) part[t] = fma(gy[(size_t)i * r + k], lv, part[t]);
        }
      }
#pragma unroll
      for (int t = 0; t < RPW; t++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part[t] += __shfl_xor_sync(0xffffffffu, part[t], o);
      if (lane == 0) {
        const double ljj = p.L_base[(size_t)j * r + j];
#pragma unroll
        for (int t = 0; t < RPW; t++) {
          const int i = warp + t * NW;
          if (i < q) gy[(size_t)i * r + j] = (gy[(size_t)i * r + j] - part[t]) / ljj;
        }
      }
      __syncwarp();
    }
    for (int i = warp; i < q; i += NW)
      for (int j = lane; j < r; j += 32) p.gSxb[(bb * q + i) * r + j] = gy[(size_t)i * r + j];
  }
}

// ---- host launchers --------------------------------------------------------------------------------
static size_t bwd_smem(int q, int r, int chunk, int qmax, bool mc_mean) {
  int QP = (q + 1) & ~1;
  int GP = q | 1;
  return ((size_t)(r + q) * QP + (size_t)q * (r + q + 1) + q + 4 * (size_t)q * q + (size_t)(chunk + 4) * GP +
          (size_t)SR_WARPS * qmax * 8 + 2 * (size_t)q + (mc_mean ? (size_t)(chunk + 4) * GP : 0)) * sizeof(double);
}

static int bwd_chunk(int q, int r, int S, int ns, bool mc_mean) {
  int chunk = ((mc_mean ? 18 : 36) * 1024) / (8 * (q | 1));
  if (chunk > S) chunk = S;
  chunk = (chunk / (SR_THREADS * ns)) * (SR_THREADS * ns);
  if (chunk < SR_THREADS * ns) chunk = SR_THREADS * ns;
  if ((int64_t)chunk * (q | 1) < (int64_t)q * r) chunk = (q * r + (q | 1) - 1) / (q | 1);
  return chunk;
}

// dynamic shared memory of the backward launch for this shape (host-side query: mcacq_fused_supported)
size_t sample_reduce_bwd_smem(int q, int r, int S, bool mc_mean) {
  const int qmax = q <= 8 ? 8 : q <= 16 ? 16 : 32;
  return bwd_smem(q, r, bwd_chunk(q, r, S, 1, mc_mean), qmax, mc_mean);
}

template <int QMAX, int NS, int W, bool PLAIN = false>
static int launch_sr_bwd(const SRParams& p, cudaStream_t st) {
  // chunk of samples whose weights are staged in shared memory (<= ~64 KB), at least r rows for the solve scratch
  const bool mc_mean = p.fat >= 5;
  const int chunk = bwd_chunk(p.q, p.r, p.S, NS, mc_mean);
  size_t smem = bwd_smem(p.q, p.r, chunk, QMAX, mc_mean);
  if (smem > 200 * 1024) return MCACQ_ELIMIT;
  auto kern = sample_reduce_bwd_kernel<QMAX, NS, W, PLAIN>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<<<(unsigned)p.b, SR_THREADS * W, smem, st>>>(p, chunk);
  count_launch();
  MCACQ_CUDA_CHECK_LAUNCH();
  return 0;
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
