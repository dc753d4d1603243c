// Backward of the posterior blocks (see blocks.cu for the forward and the data layout).
#include "blocks_common.cuh"

namespace mcacq {

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
//
// RLOOP (baselines of more than 64 points; instantiated with RT = 0): the baseline coefficients and source rows do not fit
// the register file, so the baseline term is accumulated by a run-time loop over groups of 4 baseline rows whose
// coefficient fragments are re-read (L1 / L2) in every column step -- the same DMMA sequence, in the same order, as the
// unrolled RT version.
template <int QT, int RT, int PASS, bool RLOOP = false>
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
      if (RLOOP) {
        const int i = mi * 8 + g;
        for (int j = t4; j < r; j += 4)
          if (i < q) bnd = fma(fabs(s2 * gxb[i * r + j]), p.Ab_absmax[j], bnd);
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
  // acc * 2^shift as ONE multiplication (exact: a power of two, the product stays far inside the normal range) wherever
  // 2^shift is representable; ldexp, whose range checks cost ~15 instructions per element, only for absurd row scales
  double smul[QT];
  bool sfast[QT];
#pragma unroll
  for (int mi = 0; mi < QT; mi++) {
    sfast[mi] = shift[mi] > -1000 && shift[mi] < 1000;
    smul[mi] = sfast[mi] ? __hiloint2double((1023 + shift[mi]) << 20, 0) : 1.0;
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
    if (RLOOP && p.T != nullptr) {
      // the baseline term was computed as a GEMM (T = gSxb A_base): the lane's 2 NT consecutive output columns of row g
      const int oc0 = c0 + 2 * NT * t4;
#pragma unroll
      for (int mi = 0; mi < QT; mi++) {
        const int i = mi * 8 + g;
        if (i < q && oc0 < col_end) {
          const double* src = p.T + (bb * q + i) * (int64_t)np + oc0;
#pragma unroll
          for (int e = 0; e < 2; e++)
#pragma unroll
            for (int t = 0; t < NT; t += 2) {
              const double2 v = *reinterpret_cast<const double2*>(src + e * NT + t);
              acc[mi][t][e] = -s2 * v.x;
              acc[mi][t + 1][e] = -s2 * v.y;
            }
        }
      }
    } else if (RLOOP) {
      const int cc = c0 + NT * g;   // this lane's source columns in the current step
#pragma unroll 4
      for (int j0 = 0; j0 < r; j0 += 4) {
        const int j = j0 + t4;
        double rbj[NT], cbj[QT];
        loadn<NT>(p.A_base + (int64_t)j * np + cc, j < r && cc < col_end, rbj);
#pragma unroll
        for (int mi = 0; mi < QT; mi++) {
          const int i = mi * 8 + g;
          cbj[mi] = (i < q && j < r) ? -s2 * gxb[i * r + j] : 0.0;
        }
#pragma unroll
        for (int mi = 0; mi < QT; mi++)
#pragma unroll
          for (int t = 0; t < NT; t++) dmma884b(acc[mi][t][0], acc[mi][t][1], cbj[mi], rbj[t]);
      }
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
                const double scaled = sfast[mi] ? acc[mi][t][e] * smul[mi] : ldexp(acc[mi][t][e], shift[mi]);
                const unsigned long long y = balanced_bytes(__double2ll_rn(scaled));
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
    for (int j = 0; j < (p.skip_base_direct ? 0 : r); j++) {
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

template <int QT, int RT, bool RLOOP = false>
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
    posterior_blocks_bwd_kernel<QT, RT, 1, RLOOP><<<(unsigned)blocks, BLK_WARPS * 32, 0, st>>>(p, col_chunk);
    count_launch();
    MCACQ_CUDA_CHECK_LAUNCH();
    posterior_blocks_bwd_kernel<QT, RT, 2, RLOOP><<<(unsigned)blocks, BLK_WARPS * 32, 0, st>>>(p, col_chunk);
  } else {
    posterior_blocks_bwd_kernel<QT, RT, 0, RLOOP><<<(unsigned)blocks, BLK_WARPS * 32, 0, st>>>(p, col_chunk);
  }
  count_launch();
  MCACQ_CUDA_CHECK_LAUNCH();
  return 0;
}

// Direct K(X, X_base) terms of dU for large baselines:  dU[i][k] += sum_j w_ij (u_ik - ub_jk),  w_ij = s^2 gSxb[i][j] g(|u_i - ub_j|^2).
// The tail of the block kernel evaluates them with one warp per q-batch and recomputes the distance for every (i, k): O(q d r d).
// Here a CTA per q-batch first parks the q x r weights in shared memory (one distance each), then every (i, k) sums over j.
__global__ void __launch_bounds__(128)
baseline_direct_bwd_kernel(BlocksBwdParams p) {
  extern __shared__ __align__(16) double wsm[];   // [q][r]
  const int64_t bb = blockIdx.x;
  const int q = p.q, r = p.r, d = p.d, tid = threadIdx.x;
  const double s2 = p.y_std * p.y_std;
  const double* Ub = p.U + bb * q * d;
  const double* gxb = p.gSxb + bb * q * r;
  for (int idx = tid; idx < q * r; idx += 128) {
    const int i = idx / r, j = idx - i * r;
    double sq = 0.0;
    for (int kk = 0; kk < d; kk++) {
      const double df = Ub[i * d + kk] - p.U_base[j * d + kk];
      sq = fma(df, df, sq);
    }
    wsm[idx] = s2 * gxb[idx] * kernel_dfactor(p.kernel_id, p.outputscale, sq);
  }
  __syncthreads();
  for (int idx = tid; idx < q * d; idx += 128) {
    const int i = idx / d, k = idx - i * d;
    const double uik = Ub[i * d + k];
    double accu = 0.0;
    for (int j = 0; j < r; j++) accu = fma(wsm[i * r + j], uik - p.U_base[j * d + k], accu);
    p.dU[(bb * q + i) * d + k] += accu;
  }
}

constexpr size_t BASE_DIRECT_SMEM_MAX = 160 * 1024;
bool baseline_direct_fits(int q, int r) { return (size_t)q * r * sizeof(double) <= BASE_DIRECT_SMEM_MAX; }

int baseline_direct_bwd(const BlocksBwdParams& p, cudaStream_t st) {
  const size_t smem = (size_t)p.q * p.r * sizeof(double);
  if (smem > BASE_DIRECT_SMEM_MAX) return MCACQ_ELIMIT;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(baseline_direct_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BASE_DIRECT_SMEM_MAX);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  baseline_direct_bwd_kernel<<<(unsigned)p.b, 128, smem, st>>>(p);
  count_launch();
  MCACQ_CUDA_CHECK_LAUNCH();
  return 0;
}

// dst[cols x rows] = src[rows x cols]^T (the baseline panel A_base, a few MB: once per backward call)
__global__ void transpose_panel_kernel(const double* __restrict__ src, int rows, int cols, double* __restrict__ dst) {
  __shared__ double tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? src[(int64_t)r * cols + c] : 0.0;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < rows) dst[(int64_t)c * rows + r] = tile[threadIdx.x][i];
  }
}

int transpose_panel(const double* src, int rows, int cols, double* dst, cudaStream_t st) {
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  transpose_panel_kernel<<<grid, block, 0, st>>>(src, rows, cols, dst);
  count_launch();
  MCACQ_CUDA_CHECK_LAUNCH();
  return 0;
}

static int posterior_blocks_bwd_small_r(const BlocksBwdParams& p, cudaStream_t st) {
  MCACQ_DISPATCH_QT_RT(launch_blocks_bwd, p, st);
}

int posterior_blocks_bwd(const BlocksBwdParams& p, cudaStream_t st) {
  // MCACQ_BWD_RLOOP=1 forces the run-time baseline loop for every r (tests: bit-identical to the unrolled kernels)
  const bool force_rloop = (getenv("MCACQ_BWD_RLOOP") != nullptr) && atoi(getenv("MCACQ_BWD_RLOOP")) != 0;
  if (p.r <= 64 && !(force_rloop && p.r > 0)) return posterior_blocks_bwd_small_r(p, st);
  const int qt_ = (p.q + 7) / 8;
  if (qt_ == 1) return launch_blocks_bwd<1, 0, true>(p, st);
  if (qt_ == 2) return launch_blocks_bwd<2, 0, true>(p, st);
  if (qt_ <= 4) return launch_blocks_bwd<4, 0, true>(p, st);
  return MCACQ_ELIMIT;
}

}  // namespace mcacq
