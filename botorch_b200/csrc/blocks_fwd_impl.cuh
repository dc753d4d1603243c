// Forward posterior-block kernel and its launcher (included by blocks.cu / blocks_q2.cu / blocks_q4.cu: one translation unit per
// q tile count, so that the template instantiations compile in parallel).
#pragma once
#include "blocks_common.cuh"


namespace mcacq {

// Number of k-slices (= warps per CTA) of the forward kernel: as many as keep the cross-warp reduction buffer <= 48 KB.
template <int QT, int RT, int NB>
struct FwdSplit {
  static constexpr int NACC = 2 * QT * QT + 2 * QT * RT + 2 * QT;  // Gram + cross-Gram fragments + mean + row max, per lane
  static constexpr int KS = (8 * NB * NACC * 256 <= 49152) ? 8 : (4 * NB * NACC * 256 <= 49152) ? 4
                          : (2 * NB * NACC * 256 <= 49152) ? 2 : 1;
};

// One CTA owns NB consecutive q-batches; its KS warps split the contraction dimension into KS contiguous slices, sweep them
// concurrently (an L-BFGS round of ~64 q-batches still occupies hundreds of warps) and combine the partial DMMA
// fragments through shared memory in slice order.  The decomposition depends on np only, never on b, so the results are
// bit-identical however a t-batch is chunked.
template <int QT, int RT, int NB, bool MEAN>
__global__ void __launch_bounds__(FwdSplit<QT, RT, NB>::KS * 32)
posterior_blocks_kernel(BlocksParams p) {
  constexpr int KS = FwdSplit<QT, RT, NB>::KS;
  constexpr int NACC = FwdSplit<QT, RT, NB>::NACC;
  extern __shared__ __align__(16) double red[];   // [KS][NB][NACC][32]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int64_t b0 = (int64_t)blockIdx.x * NB;
  if (b0 >= p.b) return;
  const int q = p.q, np = p.np, r = p.r;
  const int kslice = ((np / 16 + KS - 1) / KS) * 16;             // columns per warp (a multiple of the 16-column step)
  const int k_begin = warp * kslice;
  const int k_end = (k_begin + kslice < np) ? k_begin + kslice : np;

  double accG[NB][QT][QT][2];
  double accB[NB][QT][RT > 0 ? RT : 1][2];
  double macc[NB][QT];
  double amax[NB][QT];
#pragma unroll
  for (int nb = 0; nb < NB; nb++)
#pragma unroll
    for (int mi = 0; mi < QT; mi++) {
      macc[nb][mi] = 0.0;
      amax[nb][mi] = 0.0;
#pragma unroll
      for (int nj = 0; nj < QT; nj++) { accG[nb][mi][nj][0] = 0.0; accG[nb][mi][nj][1] = 0.0; }
#pragma unroll
      for (int nj = 0; nj < (RT > 0 ? RT : 1); nj++) { accB[nb][mi][nj][0] = 0.0; accB[nb][mi][nj][1] = 0.0; }
    }

  // Software-pipelined sweep: every A fragment register is re-loaded for the NEXT 16-column step right after its last
  // use in the current one, so a full step of DMMAs (NB batches) covers the latency of each 32-byte load without any
  // extra registers.
  double af[NB][QT][4];
  double kf[MEAN ? NB : 1][QT][4];
  double bf[RT > 0 ? RT : 1][4];
  double al[4] = {0.0, 0.0, 0.0, 0.0};
  const int64_t off0 = (b0 * q + g) * (int64_t)np + 4 * t4;   // fragment (nb, mi) starts at off0 + (nb * q + 8 * mi) * np
  const int64_t bstride = (int64_t)q * np;
  const bool live = k_begin < k_end;
#pragma unroll
  for (int nb = 0; nb < NB; nb++)
#pragma unroll
    for (int mi = 0; mi < QT; mi++) {
      const bool ok = live && (b0 + nb < p.b) && mi * 8 + g < q;
      load4(p.A + off0 + nb * bstride + (int64_t)mi * 8 * np + k_begin, ok, af[nb][mi]);
      if (MEAN) load4(p.Kt + off0 + nb * bstride + (int64_t)mi * 8 * np + k_begin, ok, kf[nb][mi]);
    }
#pragma unroll
  for (int nj = 0; nj < RT; nj++)
    load4(p.A_base + (int64_t)(nj * 8 + g) * np + k_begin + 4 * t4, live && nj * 8 + g < r, bf[nj]);
  if (MEAN && live) load4(p.alpha + k_begin + 4 * t4, true, al);

  for (int k0 = k_begin; k0 < k_end; k0 += 16) {
    const bool more = k0 + 16 < k_end;
#pragma unroll
    for (int nb = 0; nb < NB; nb++) {
#pragma unroll
      for (int mi = 0; mi < QT; mi++) {
        if (MEAN) {
#pragma unroll
          for (int s = 0; s < 4; s++) macc[nb][mi] = fma(kf[nb][mi][s], al[s], macc[nb][mi]);
          load4(p.Kt + off0 + nb * bstride + (int64_t)mi * 8 * np + k0 + 16, more && (b0 + nb < p.b) && mi * 8 + g < q,
                kf[nb][mi]);
        }
#pragma unroll
        for (int s = 0; s < 4; s++) amax[nb][mi] = fmax(amax[nb][mi], fabs(af[nb][mi][s]));
      }
#pragma unroll
      for (int s = 0; s < 4; s++)
#pragma unroll
        for (int mi = 0; mi < QT; mi++) {
#pragma unroll
          for (int nj = 0; nj <= mi; nj++)
            dmma884b(accG[nb][mi][nj][0], accG[nb][mi][nj][1], af[nb][mi][s], af[nb][nj][s]);
#pragma unroll
          for (int nj = 0; nj < RT; nj++) dmma884b(accB[nb][mi][nj][0], accB[nb][mi][nj][1], af[nb][mi][s], bf[nj][s]);
        }
#pragma unroll
      for (int mi = 0; mi < QT; mi++)
        load4(p.A + off0 + nb * bstride + (int64_t)mi * 8 * np + k0 + 16, more && (b0 + nb < p.b) && mi * 8 + g < q,
              af[nb][mi]);
    }
    // the baseline fragments are shared by all batches of the step (L2-resident r x np panel): reload after the last use
#pragma unroll
    for (int nj = 0; nj < RT; nj++)
      load4(p.A_base + (int64_t)(nj * 8 + g) * np + k0 + 16 + 4 * t4, more && nj * 8 + g < r, bf[nj]);
    if (MEAN) load4(p.alpha + k0 + 16 + 4 * t4, more, al);
  }

  // ---- combine the KS partial results (slice order) -- warp nb finishes q-batch b0 + nb
  if (KS > 1) {
#pragma unroll
    for (int nb = 0; nb < NB; nb++) {
      double* dst = red + ((size_t)(warp * NB + nb) * NACC) * 32 + lane;
      int a = 0;
#pragma unroll
      for (int mi = 0; mi < QT; mi++) {
#pragma unroll
        for (int nj = 0; nj < QT; nj++) { dst[32 * a++] = accG[nb][mi][nj][0]; dst[32 * a++] = accG[nb][mi][nj][1]; }
#pragma unroll
        for (int nj = 0; nj < RT; nj++) { dst[32 * a++] = accB[nb][mi][nj][0]; dst[32 * a++] = accB[nb][mi][nj][1]; }
        dst[32 * a++] = macc[nb][mi];
        dst[32 * a++] = amax[nb][mi];
      }
    }
    __syncthreads();
  }
  const double s2 = p.y_std * p.y_std;
  for (int nb = (KS > 1 ? warp : 0); nb < NB; nb += (KS > 1 ? KS : 1)) {
    const int64_t bb = b0 + nb;
    if (bb >= p.b) continue;
    double G_[QT][QT][2], B_[QT][RT > 0 ? RT : 1][2], mv_[QT], av_[QT];
    if (KS > 1) {
#pragma unroll
      for (int mi = 0; mi < QT; mi++) {
        mv_[mi] = 0.0; av_[mi] = 0.0;
#pragma unroll
        for (int nj = 0; nj < QT; nj++) { G_[mi][nj][0] = 0.0; G_[mi][nj][1] = 0.0; }
#pragma unroll
        for (int nj = 0; nj < (RT > 0 ? RT : 1); nj++) { B_[mi][nj][0] = 0.0; B_[mi][nj][1] = 0.0; }
      }
      for (int w = 0; w < KS; w++) {
        const double* src = red + ((size_t)(w * NB + nb) * NACC) * 32 + lane;
        int a = 0;
#pragma unroll
        for (int mi = 0; mi < QT; mi++) {
#pragma unroll
          for (int nj = 0; nj < QT; nj++) { G_[mi][nj][0] += src[32 * a++]; G_[mi][nj][1] += src[32 * a++]; }
#pragma unroll
          for (int nj = 0; nj < RT; nj++) { B_[mi][nj][0] += src[32 * a++]; B_[mi][nj][1] += src[32 * a++]; }
          mv_[mi] += src[32 * a++];
          av_[mi] = fmax(av_[mi], src[32 * a++]);
        }
      }
    } else {
#pragma unroll
      for (int mi = 0; mi < QT; mi++) {
        mv_[mi] = macc[nb][mi]; av_[mi] = amax[nb][mi];
#pragma unroll
        for (int nj = 0; nj < QT; nj++) { G_[mi][nj][0] = accG[nb][mi][nj][0]; G_[mi][nj][1] = accG[nb][mi][nj][1]; }
#pragma unroll
        for (int nj = 0; nj < RT; nj++) { B_[mi][nj][0] = accB[nb][mi][nj][0]; B_[mi][nj][1] = accB[nb][mi][nj][1]; }
      }
    }
    const double* Ub = p.U + bb * q * p.d;
#pragma unroll
    for (int mi = 0; mi < QT; mi++) {
      // mean: reduce the 4 lanes of a row group
      double mv = mv_[mi];
      mv += __shfl_xor_sync(0xffffffffu, mv, 1);
      mv += __shfl_xor_sync(0xffffffffu, mv, 2);
      double av = av_[mi];
      av = fmax(av, __shfl_xor_sync(0xffffffffu, av, 1));
      av = fmax(av, __shfl_xor_sync(0xffffffffu, av, 2));
      const int i = mi * 8 + g;
      if (t4 == 0 && i < q && !p.cross_only) {
        if (p.mean_part != nullptr) {  // int8 mode: Kt * alpha was reduced per 512-column tile by the covariance kernel
          const int64_t Mrows = p.b * q;
          mv = 0.0;
          for (int t = 0; t < p.n_parts; t++) mv += p.mean_part[(int64_t)t * Mrows + bb * q + i];
        }
        p.mean[bb * q + i] = p.y_mean + p.y_std * (p.mean_const + mv);
        if (p.A_absmax != nullptr) p.A_absmax[bb * q + i] = av;
      }
      if (i >= q) continue;
#pragma unroll
      for (int nj = 0; nj <= mi; nj++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int j = nj * 8 + 2 * t4 + e;
          if (j >= q || p.cross_only) continue;
          double sq = 0.0;
          for (int k = 0; k < p.d; k++) {
            double df = Ub[i * p.d + k] - Ub[j * p.d + k];
            sq = fma(df, df, sq);
          }
          double v = s2 * (kernel_value(p.kernel_id, p.outputscale, sq) - G_[mi][nj][e]);
          p.Sxx[(bb * q + i) * q + j] = v;
          if (nj < mi) p.Sxx[(bb * q + j) * q + i] = v;
        }
#pragma unroll
      for (int nj = 0; nj < RT; nj++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int j = nj * 8 + 2 * t4 + e;
          if (j >= r) continue;
          double sq = 0.0;
          for (int k = 0; k < p.d; k++) {
            double df = Ub[i * p.d + k] - p.U_base[j * p.d + k];
            sq = fma(df, df, sq);
          }
          p.Sxb[(bb * q + i) * (int64_t)p.r_pitch + j] = s2 * (kernel_value(p.kernel_id, p.outputscale, sq) - B_[mi][nj][e]);
        }
    }
  }
}

template <int QT, int RT>
static int launch_blocks_fwd(const BlocksParams& p, cudaStream_t st) {
  constexpr int NB = (QT == 1 && RT <= 4) ? 2 : 1;
  constexpr int KS = FwdSplit<QT, RT, NB>::KS;
  const size_t smem = (KS > 1) ? (size_t)KS * NB * FwdSplit<QT, RT, NB>::NACC * 32 * sizeof(double) : 0;
  int64_t blocks = (p.b + NB - 1) / NB;
  if (p.Kt != nullptr) posterior_blocks_kernel<QT, RT, NB, true><<<(unsigned)blocks, KS * 32, smem, st>>>(p);
  else posterior_blocks_kernel<QT, RT, NB, false><<<(unsigned)blocks, KS * 32, smem, st>>>(p);
  count_launch();
  MCACQ_CUDA_CHECK_LAUNCH();
  return 0;
}


}  // namespace mcacq
