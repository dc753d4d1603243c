// Posterior blocks of a q-batch from the contracted rows A = K(X, X_train) R.
//
//   mean[b][i]    = m + s * (c + sum_k Kt[i][k] alpha[k])            gpytorch exact_predictive_mean
//   Sxx[b][i][j]  = s^2 * (k(u_i, u_j)  - sum_k A[i][k] A[j][k])     gpytorch exact_predictive_covar
//   Sxb[b][i][j'] = s^2 * (k(u_i, ub_j') - sum_k A[i][k] A_base[j'][k])   (rows -q: of the joint covariance,
//                                                                     utils/low_rank.py:117-124)
// followed by Standardize.untransform_posterior (transforms/outcome.py:479-511).
//
// Hardware mapping: one warp owns NB consecutive q-batches and sweeps k in steps of 16.  The Gram and
// cross-Gram tiles are DMMA.8x8x4 contractions whose operands are loaded straight from global memory
// as 32-byte per-thread vectors: lane (g, t) holds row g, columns k0+4t..k0+4t+3, and the four
// consecutive DMMA steps use a k-permutation (virtual k = (step, t) <-> actual k0 + 4t + step) that is
// applied identically to both operands, so no shared-memory staging or shuffles are needed.
#include <cstdlib>

#include "common.cuh"
#include "params.cuh"

namespace mcacq {

__device__ __forceinline__ void dmma884b(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void load4(const double* p, bool ok, double (&v)[4]) {
  if (ok) {
    double2 a = *reinterpret_cast<const double2*>(p);
    double2 b = *reinterpret_cast<const double2*>(p + 2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  } else {
    v[0] = v[1] = v[2] = v[3] = 0.0;
  }
}


template <int N>
__device__ __forceinline__ void loadn(const double* p, bool ok, double (&v)[N]) {
  static_assert(N == 2 || N == 4, "fragment width");
  if (ok) {
    double2 a = *reinterpret_cast<const double2*>(p);
    v[0] = a.x; v[1] = a.y;
    if (N == 4) {
      double2 b = *reinterpret_cast<const double2*>(p + 2);
      v[2] = b.x; v[3] = b.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; i++) v[i] = 0.0;
  }
}

constexpr int BLK_WARPS = 4;

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
      if (t4 == 0 && i < q) {
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
          if (j >= q) continue;
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
          p.Sxb[(bb * q + i) * r + j] = s2 * (kernel_value(p.kernel_id, p.outputscale, sq) - B_[mi][nj][e]);
        }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Backward:  dA[i][:] = -s^2 * ( sum_j (gSxx[i][j] + gSxx[j][i]) A[j][:] + sum_j' gSxb[i][j'] A_base[j'][:] )
// written IN PLACE over A, row_scale[i] = s * gmean[i] (the rank-1 mean term is applied by the covariance
// backward kernel), and the direct kernel terms of K(X, X) and K(X, X_base) into dU.

//
// int8 mode (emit_slices): the G signed 8-bit slices of dA are emitted instead of the fp64 matrix, each row in fixed point
// relative to a power of two above its largest entry.  PASS = 1 computes that largest entry (the same DMMA sequence, no
// stores; atomicMax of the exponent over the column chunks of a row), PASS = 2 emits the slices with it.  Scaling by the
// ACTUAL row maximum matters: the a-priori bound sum_j |C[i][j]| max|A[j][:]| (PASS = 0, kept behind MCACQ_DA_BOUND=1)
// overestimates rows whose terms cancel -- nearly collinear rows of A with alternating coefficients, i.e. exactly the
// ill-conditioned q-batches -- by the inverse of the smallest relative pivot, and every factor 256 costs one slice.
template <int QT, int RT, int PASS>
__global__ void __launch_bounds__(BLK_WARPS * 32, (QT == 1 && RT <= 2) ? 4 : 1)
posterior_blocks_bwd_kernel(BlocksBwdParams p, int col_chunk) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  // one warp per (q-batch, chunk of BWD_CHUNK output columns): the columns are independent, so splitting them keeps
  // every result bit-identical while an L-BFGS round (b = num_restarts ~ 64 q-batches) still fills the machine
  const int n_chunks = (p.np + col_chunk - 1) / col_chunk;
  const int64_t wid = (int64_t)blockIdx.x * BLK_WARPS + warp;
  const int64_t bb = wid / n_chunks;
  const int chunk = (int)(wid - bb * n_chunks);
  if (bb >= p.b) return;
  const int q = p.q, np = p.np, r = p.r, d = p.d;
  const int col_begin = chunk * col_chunk, col_end = (col_begin + col_chunk < np) ? col_begin + col_chunk : np;
  const double s2 = p.y_std * p.y_std;
  const double* gxx = p.gSxx + bb * q * q;
  const double* gxb = p.gSxb + bb * q * r;

  // coefficient fragments (A operand): row i = 8*mi + g, contracted index = 4*kk + t4
  double cq[QT][2 * QT];
  double cb[QT][RT > 0 ? 2 * RT : 1];
#pragma unroll
  for (int mi = 0; mi < QT; mi++) {
    const int i = mi * 8 + g;
#pragma unroll
    for (int kk = 0; kk < 2 * QT; kk++) {
      const int j = kk * 4 + t4;
      cq[mi][kk] = (i < q && j < q) ? -s2 * (gxx[i * q + j] + gxx[j * q + i]) : 0.0;
    }
#pragma unroll
    for (int kk = 0; kk < 2 * RT; kk++) {
      const int j = kk * 4 + t4;
      cb[mi][kk] = (i < q && j < r) ? -s2 * gxb[i * r + j] : 0.0;
    }
  }

  // int8 mode: exponent of an upper bound of |dA[i][:]| <= sum_j |C[i][j]| max|A[j][:]| + sum_j' |Cb[i][j']| max|A_base[j'][:]|
  int shift[QT];
  double rowmax[QT];
#pragma unroll
  for (int mi = 0; mi < QT; mi++) { shift[mi] = 0; rowmax[mi] = 0.0; }
  if (PASS == 2) {
#pragma unroll
    for (int mi = 0; mi < QT; mi++) {
      const int i = mi * 8 + g;
      int ex = (i < q) ? p.slice_exp[bb * q + i] : 0;
      if (ex < -2000) ex = 0;   // an all-zero (or non-finite) row: every digit is zero whatever the scale
      shift[mi] = 8 * p.G - 2 - ex;
      if (t4 == 0 && i < q && chunk == 0) p.slice_scale[bb * q + i] = ldexp(1.0, ex + 2);
    }
  }
  if (PASS == 0 && p.emit_slices) {
#pragma unroll
    for (int mi = 0; mi < QT; mi++) {
      double bnd = 0.0;
#pragma unroll
      for (int kk = 0; kk < 2 * QT; kk++) {
        const int j = kk * 4 + t4;
        if (j < q) bnd = fma(fabs(cq[mi][kk]), p.A_absmax[bb * q + j], bnd);
      }
#pragma unroll
      for (int kk = 0; kk < 2 * RT; kk++) {
        const int j = kk * 4 + t4;
        if (j < r) bnd = fma(fabs(cb[mi][kk]), p.Ab_absmax[j], bnd);
      }
      bnd += __shfl_xor_sync(0xffffffffu, bnd, 1);
      bnd += __shfl_xor_sync(0xffffffffu, bnd, 2);
      int ex = 0;
      if (bnd > 0.0 && isfinite(bnd)) frexp(bnd * (1.0 + 1e-9), &ex);
      shift[mi] = 8 * p.G - 2 - ex;
      const int i = mi * 8 + g;
      if (t4 == 0 && i < q && chunk == 0) p.slice_scale[bb * q + i] = ldexp(1.0, ex + 2);
    }
  }
  const size_t slice_stride = (size_t)p.b * q * np;
  double* Ab = p.A + bb * q * np;
  // 8 * NT output columns per step as NT DMMA column tiles (NT = 4 for the small shapes, 2 when the coefficient
  // fragments already fill the register file).  Lane (g, t4) loads columns c0 + NT g .. + NT-1 of source row
  // j = 4 kk + t4; tile t takes column NT n + t as its B column n, so the lane ends up with the 2 NT CONSECUTIVE output
  // columns c0 + 2 NT t4 .. of row g (acc[.][t][e] <-> column 2 NT t4 + NT e + t): 64-byte fp64 stores / 8-byte slice
  // stores at NT = 4.  The source fragments are re-loaded for the next step right after their last DMMA (software
  // pipelining without extra registers); the in-place dA store only touches columns of the current step.
  constexpr int NT = (QT == 1 && RT <= 4) ? 4 : 2;
  double rq[2 * QT][NT];
  double rb[RT > 0 ? 2 * RT : 1][NT];
#pragma unroll
  for (int kk = 0; kk < 2 * QT; kk++)
    loadn<NT>(Ab + (int64_t)(kk * 4 + t4) * np + col_begin + NT * g, kk * 4 + t4 < q && col_begin + NT * g < col_end, rq[kk]);
#pragma unroll
  for (int kk = 0; kk < 2 * RT; kk++)
    loadn<NT>(p.A_base + (int64_t)(kk * 4 + t4) * np + col_begin + NT * g, kk * 4 + t4 < r && col_begin + NT * g < col_end, rb[kk]);
  for (int c0 = col_begin; c0 < col_end; c0 += 8 * NT) {
    const int cn = c0 + 8 * NT + NT * g;  // this lane's source columns in the next step
    double acc[QT][NT][2];
#pragma unroll
    for (int mi = 0; mi < QT; mi++)
#pragma unroll
      for (int t = 0; t < NT; t++) { acc[mi][t][0] = 0.0; acc[mi][t][1] = 0.0; }
#pragma unroll
    for (int kk = 0; kk < 2 * RT; kk++) {
#pragma unroll
      for (int mi = 0; mi < QT; mi++)
#pragma unroll
        for (int t = 0; t < NT; t++) dmma884b(acc[mi][t][0], acc[mi][t][1], cb[mi][kk], rb[kk][t]);
      loadn<NT>(p.A_base + (int64_t)(kk * 4 + t4) * np + cn, kk * 4 + t4 < r && cn < col_end, rb[kk]);
    }
#pragma unroll
    for (int kk = 0; kk < 2 * QT; kk++) {
#pragma unroll
      for (int mi = 0; mi < QT; mi++)
#pragma unroll
        for (int t = 0; t < NT; t++) dmma884b(acc[mi][t][0], acc[mi][t][1], cq[mi][kk], rq[kk][t]);
      loadn<NT>(Ab + (int64_t)(kk * 4 + t4) * np + cn, kk * 4 + t4 < q && cn < col_end, rq[kk]);
    }
    // every lane of the warp has issued its loads of this step's columns (they were fetched one step earlier) before
    // any lane overwrites them in place
    __syncwarp();
    const int oc = c0 + 2 * NT * t4;
    if (PASS == 1) {
      // columns beyond col_end were loaded as zeros, rows beyond q have zero coefficients: their accumulators are 0
#pragma unroll
      for (int mi = 0; mi < QT; mi++)
#pragma unroll
        for (int t = 0; t < NT; t++) rowmax[mi] = fmax(rowmax[mi], fmax(fabs(acc[mi][t][0]), fabs(acc[mi][t][1])));
      continue;
    }
    if (oc < col_end) {
#pragma unroll
      for (int mi = 0; mi < QT; mi++) {
        const int i = mi * 8 + g;
        if (i < q) {
          if (PASS == 2 || p.emit_slices) {
            unsigned long long Y[2][4];   // NT = 4: [e][t];  NT = 2: Y[0] = the lane's 4 columns
#pragma unroll
            for (int e = 0; e < 2; e++)
#pragma unroll
              for (int t = 0; t < NT; t++) {
                const unsigned long long y = balanced_bytes(__double2ll_rn(ldexp(acc[mi][t][e], shift[mi])));
                if (NT == 4) Y[e][t] = y; else Y[0][e * 2 + t] = y;
              }
            int8_t* sdst = p.slices + ((size_t)(bb * q + i)) * np + oc;
            if (NT == 4)
              digits4x2(Y[0], Y[1], [&](int dg, unsigned wa, unsigned wb) {
                if (dg < p.G) *reinterpret_cast<uint2*>(sdst + (size_t)(p.G - 1 - dg) * slice_stride) = make_uint2(wa, wb);
              });
            else
              digits4(Y[0], [&](int dg, unsigned w) {
                if (dg < p.G) *reinterpret_cast<unsigned*>(sdst + (size_t)(p.G - 1 - dg) * slice_stride) = w;
              });
          } else {
            double* dst = Ab + (int64_t)i * np + oc;
#pragma unroll
            for (int e = 0; e < 2; e++)
#pragma unroll
              for (int t = 0; t < NT; t += 2)
                *reinterpret_cast<double2*>(dst + e * NT + t) = make_double2(acc[mi][t][e], acc[mi][t + 1][e]);
          }
        }
      }
    }
  }

  if (PASS == 1) {
#pragma unroll
    for (int mi = 0; mi < QT; mi++) {
      double mx = rowmax[mi];
      mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const int i = mi * 8 + g;
      if (t4 == 0 && i < q && mx > 0.0 && isfinite(mx)) {
        int ex = 0;
        frexp(mx, &ex);   // mx = m 2^ex, m in [0.5, 1): every entry of the row is below 2^ex
        atomicMax(p.slice_exp + bb * q + i, ex);
      }
    }
    return;
  }
  // rank-1 mean term scale and direct kernel terms (once per q-batch)
  if (chunk != 0) return;
  const double* Ub = p.U + bb * q * d;
  for (int idx = lane; idx < q; idx += 32) p.row_scale[bb * q + idx] = p.y_std * p.gmean[bb * q + idx];
  for (int idx = lane; idx < q * d; idx += 32) {
    const int i = idx / d, k = idx - i * d;
    double accu = 0.0;
    for (int j = 0; j < q; j++) {
      if (j == i) continue;
      double sq = 0.0;
      for (int kk = 0; kk < d; kk++) {
        double df = Ub[i * d + kk] - Ub[j * d + kk];
        sq = fma(df, df, sq);
      }
      double w = s2 * (gxx[i * q + j] + gxx[j * q + i]) * kernel_dfactor(p.kernel_id, p.outputscale, sq);
      accu = fma(w, Ub[i * d + k] - Ub[j * d + k], accu);
    }
    for (int j = 0; j < r; j++) {
      double sq = 0.0;
      for (int kk = 0; kk < d; kk++) {
        double df = Ub[i * d + kk] - p.U_base[j * d + kk];
        sq = fma(df, df, sq);
      }
      double w = s2 * gxb[i * r + j] * kernel_dfactor(p.kernel_id, p.outputscale, sq);
      accu = fma(w, Ub[i * d + k] - p.U_base[j * d + k], accu);
    }
    p.dU[(bb * q + i) * d + k] = accu;
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

template <int QT, int RT>
static int launch_blocks_bwd(const BlocksBwdParams& p, cudaStream_t st) {
  // the output columns are independent, so the chunk width is free to follow the batch size: one warp per q-batch when
  // there are thousands of them (no duplicated preamble), narrow chunks when an L-BFGS round brings only a few dozen
  int col_chunk = ((p.np + 31) / 32) * 32;
  if (p.b < 2048) {  // aim at >= 2048 warps, at least 4 steps of 32 columns each
    int64_t c = ((p.b * (int64_t)p.np / 2048) / 32) * 32;
    col_chunk = (int)(c < 128 ? 128 : (c > col_chunk ? col_chunk : c));
  }
  const int64_t warps = p.b * ((p.np + col_chunk - 1) / col_chunk);
  int64_t blocks = (warps + BLK_WARPS - 1) / BLK_WARPS;
  static const bool use_bound = (getenv("MCACQ_DA_BOUND") != nullptr) && atoi(getenv("MCACQ_DA_BOUND")) != 0;
  if (p.emit_slices && !use_bound) {
    // 0x80808080 = -2139062144: below every frexp exponent
    if (cudaMemsetAsync(p.slice_exp, 0x80, (size_t)p.b * p.q * sizeof(int32_t), st) != cudaSuccess)
      return (int)cudaGetLastError();
    posterior_blocks_bwd_kernel<QT, RT, 1><<<(unsigned)blocks, BLK_WARPS * 32, 0, st>>>(p, col_chunk);
    count_launch();
    MCACQ_CUDA_CHECK_LAUNCH();
    posterior_blocks_bwd_kernel<QT, RT, 2><<<(unsigned)blocks, BLK_WARPS * 32, 0, st>>>(p, col_chunk);
  } else {
    posterior_blocks_bwd_kernel<QT, RT, 0><<<(unsigned)blocks, BLK_WARPS * 32, 0, st>>>(p, col_chunk);
  }
  count_launch();
  MCACQ_CUDA_CHECK_LAUNCH();
  return 0;
}

#define MCACQ_DISPATCH_QT_RT(FN, p, st)                                   \
  do {                                                                    \
    const int qt_ = (p.q + 7) / 8, rt_ = (p.r + 7) / 8;                   \
    if (qt_ == 1) {                                                       \
      if (rt_ == 0) return FN<1, 0>(p, st);                               \
      if (rt_ == 1) return FN<1, 1>(p, st);                               \
      if (rt_ == 2) return FN<1, 2>(p, st);                               \
      if (rt_ <= 4) return FN<1, 4>(p, st);                               \
      if (rt_ <= 8) return FN<1, 8>(p, st);                               \
    } else if (qt_ == 2) {                                                \
      if (rt_ == 0) return FN<2, 0>(p, st);                               \
      if (rt_ <= 2) return FN<2, 2>(p, st);                               \
      if (rt_ <= 4) return FN<2, 4>(p, st);                               \
      if (rt_ <= 8) return FN<2, 8>(p, st);                               \
    } else if (qt_ <= 4) {                                                \
      if (rt_ == 0) return FN<4, 0>(p, st);                               \
      if (rt_ <= 4) return FN<4, 4>(p, st);                               \
      if (rt_ <= 8) return FN<4, 8>(p, st);                               \
    }                                                                     \
    return MCACQ_ELIMIT;                                                  \
  } while (0)

int posterior_blocks_fwd(const BlocksParams& p, cudaStream_t st) { MCACQ_DISPATCH_QT_RT(launch_blocks_fwd, p, st); }
int posterior_blocks_bwd(const BlocksBwdParams& p, cudaStream_t st) { MCACQ_DISPATCH_QT_RT(launch_blocks_bwd, p, st); }

}  // namespace mcacq
