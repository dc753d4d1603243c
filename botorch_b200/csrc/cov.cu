// Fused ARD cross-covariance kernels (RBF / Matern-5/2, optional ScaleKernel outputscale).
//
// Replaces gpytorch RBFKernel / MaternKernel(nu=2.5) / ScaleKernel forward as BoTorch configures them
// (botorch/models/utils/gpytorch_modules.py:100-133, botorch/models/gp_regression.py:192-196) and
// Normalize._transform (botorch/models/transforms/input.py:541-554).  Distances are formed directly
// as sum_k (u_ik - u_jk)^2 on inputs already divided by the lengthscale (no GEMM expansion, so no
// cancellation), the training set tile is staged in shared memory with coalesced loads, and the
// output rows are written with 32-byte vector stores.
#include "common.cuh"

namespace mcacq {

thread_local int g_launch_count = 0;

__global__ void scale_inputs_kernel(const double* __restrict__ X, int64_t total, int d,
                                    const double* __restrict__ off, const double* __restrict__ coef,
                                    const double* __restrict__ ls, double* __restrict__ U) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= total) return;
  int k = (int)(i % d);
  U[i] = ((X[i] - off[k]) / coef[k]) / ls[k];
}

__global__ void unscale_grad_kernel(const double* __restrict__ dU, int64_t total, int d,
                                    const double* __restrict__ coef, const double* __restrict__ ls,
                                    double* __restrict__ dX) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= total) return;
  int k = (int)(i % d);
  dX[i] = (dU[i] / ls[k]) / coef[k];
}

// ---------------------------------------------------------------------------------------------
// Forward: 64 x 64 output tile per CTA, 256 threads as 16 x 16, 4 x 4 outputs per thread.
// Shared memory: U1^T tile [d][64], U2^T tile [d][64]  (dimension-major, so a thread's 4 points are
// one 32-byte vector and a half-warp reads 512 contiguous bytes).
constexpr int CT = 64;
constexpr int CT_PER_CTA = 8;  // column tiles swept by one CTA in the slice-emitting variant

// SLICES = true (int8 contraction mode): instead of the fp64 matrix the kernel emits the G signed 8-bit slices of
// every entry (fixed exponent: 0 < k <= outputscale < 2^e; same digits as slice_rows_kernel in ozaki_imma.cu) and the
// per-tile partial sums of K * alpha, so neither the fp64 K nor a separate slicing pass touches HBM.
template <bool SLICES>
__global__ void __launch_bounds__(256, 3)
cov_cross_kernel(int kernel_id, double outputscale, const double* __restrict__ U1, int64_t m1,
                 const double* __restrict__ U2, int m2, int d, double* __restrict__ K, int64_t ldk,
                 int8_t* __restrict__ S, int G, int fixed_exp, const double* __restrict__ alpha,
                 double* __restrict__ mean_part, int64_t m1_total, int64_t row_base) {
  extern __shared__ __align__(16) double sm[];
  double* s1 = sm;             // [d][CT]
  double* s2 = sm + d * CT;    // [d][CT]
  const int tid = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.y * CT;
  const int tx = tid & 15, ty = tid >> 4;
  // SLICES: one CTA sweeps CT_PER_CTA consecutive column tiles so that the K * alpha partial sums need only
  // ceil(tiles / CT_PER_CTA) slots per row; otherwise one tile per CTA.
  const int tiles_per_cta = SLICES ? CT_PER_CTA : 1;
  const int n_col_tiles = (int)((ldk + CT - 1) / CT);
  const int t_begin = blockIdx.x * tiles_per_cta;
  const int t_end = (t_begin + tiles_per_cta < n_col_tiles) ? t_begin + tiles_per_cta : n_col_tiles;

  for (int idx = tid; idx < CT * d; idx += 256) {
    int p = idx / d, k = idx - p * d;  // coalesced read of point-major global rows
    int64_t gr = row0 + p;
    s1[k * CT + p] = (gr < m1) ? U1[gr * d + k] : 0.0;
  }
  const int stage_p0 = tid / d, stage_k0 = tid - stage_p0 * d, stage_dq = 256 / d, stage_dr = 256 - stage_dq * d;
  double pm[4] = {0.0, 0.0, 0.0, 0.0};
  // 2^(8G - 2 - e): |shift| stays far inside the exponent range (0 < outputscale < 2^e is a kernel hyper-parameter)
  const double slice_mul = SLICES ? ldexp(1.0, 8 * G - 2 - fixed_exp) : 1.0;

  for (int ct = t_begin; ct < t_end; ct++) {
    const int col0 = ct * CT;
    __syncthreads();  // previous tile's reads of s2 are done (and s1 is staged on the first pass)
    {
      // (point p, dimension k) of element idx = tid + 256 i without a division per element; its address is simply
      // col0 * d + idx, the tile's rows being contiguous in the point-major operand
      const double* src = U2 + (int64_t)col0 * d;
      int pp = stage_p0, kk = stage_k0;
      for (int idx = tid; idx < CT * d; idx += 256) {
        s2[kk * CT + pp] = (col0 + pp < m2) ? src[idx] : 0.0;
        kk += stage_dr; pp += stage_dq;
        if (kk >= d) { kk -= d; pp++; }
      }
    }
    __syncthreads();

    double sq[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) sq[i][j] = 0.0;

    for (int k = 0; k < d; k++) {
      const double4 a4 = *reinterpret_cast<const double4*>(s1 + k * CT + ty * 4);
      const double4 b4 = *reinterpret_cast<const double4*>(s2 + k * CT + tx * 4);
      const double a[4] = {a4.x, a4.y, a4.z, a4.w};
      const double bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
          double df = a[i] - bb[j];
          sq[i][j] = fma(df, df, sq[i][j]);
        }
    }

#pragma unroll
    for (int i = 0; i < 4; i++) {
      int64_t gr = row0 + ty * 4 + i;
      const bool row_ok = gr < m1;
      if (!row_ok) continue;
      double v[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        int gc = col0 + tx * 4 + j;
        v[j] = (gc < m2) ? kernel_value(kernel_id, outputscale, sq[i][j]) : 0.0;
      }
      int gc0 = col0 + tx * 4;
      if (SLICES) {
#pragma unroll
        for (int j = 0; j < 4; j++) pm[i] = fma(v[j], (gc0 + j < m2) ? alpha[gc0 + j] : 0.0, pm[i]);
        if (gc0 < ldk) {  // ldk is a multiple of 16, so the 4 columns are all inside the pitch
          unsigned long long Y[4];
#pragma unroll
          for (int j = 0; j < 4; j++) Y[j] = balanced_bytes(__double2ll_rn(v[j] * slice_mul));   // exact: a power of two
          int8_t* sdst = S + ((size_t)row_base + gr) * ldk + gc0;
          const size_t sstride = (size_t)m1_total * ldk;
          digits4(Y, [&](int dg, unsigned w) {  // slice pp = digit G-1-pp (most significant first)
            if (dg < G) *reinterpret_cast<unsigned*>(sdst + (size_t)(G - 1 - dg) * sstride) = w;
          });
        }
        continue;
      }
      double* dst = K + gr * ldk + gc0;
      if (gc0 + 3 < ldk && ((ldk & 1) == 0)) {
        *reinterpret_cast<double2*>(dst) = make_double2(v[0], v[1]);
        *reinterpret_cast<double2*>(dst + 2) = make_double2(v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; j++)
          if (gc0 + j < ldk) dst[j] = v[j];
      }
    }
  }
  if (SLICES) {
    // K * alpha over this CTA's columns: deterministic shuffle tree over the 16 lanes of a row, one writer per row
#pragma unroll
    for (int i = 0; i < 4; i++) {
      double v = pm[i];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      const int64_t gr = row0 + ty * 4 + i;
      if (tx == 0 && gr < m1) mean_part[(int64_t)blockIdx.x * m1_total + row_base + gr] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Backward: dU1[i][:] (+)= sum_j G[i][j] * (u1_i - u2_j),  G[i][j] = (W[i][j] + row_scale[i] * col_vec[j]) * g(sq_ij)
// evaluated in the split form  u1_i * (sum_j G[i][j]) - (G U2)[i][:]  (the same form autograd produces for the
// reference's GEMM-expanded distance).  One CTA owns 64 rows and sweeps the columns in tiles of 64:
//   phase 1 (SIMT, 4 x 4 outputs per thread as in the forward kernel): sq, g(sq), G tile -> shared memory;
//   phase 2 (tensor pipe): V[64 x DP] += G[64 x 64] * [U2 | 1][64 x DP] with DMMA.8x8x4, one 8-row tile per warp;
//   the extra all-ones column yields the row sums for free.
constexpr int BW_T = 64;        // rows per CTA and columns per tile
constexpr int BW_P = BW_T + 4;  // padded pitch: conflict-free 8x4 / 4x8 fragment reads
constexpr int BW_CHUNK = 512;   // columns per CTA in split mode (a multiple of BW_T, >= MCACQ_MAX_D)

__device__ __forceinline__ void dmma884c(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int NTD>  // number of 8-wide dimension tiles: 8 * NTD >= d + 1
__global__ void __launch_bounds__(256, 2)
cov_cross_bwd_kernel(int kernel_id, double outputscale, const double* __restrict__ U1, int64_t m1,
                     const double* __restrict__ U2, int m2, int d, const double* __restrict__ W, int64_t ldw,
                     const double* __restrict__ row_scale, const double* __restrict__ col_vec,
                     double* __restrict__ dU1, int accumulate, int col_chunk, double* __restrict__ parts) {
  // col_chunk > 0 (split mode): blockIdx.y owns the columns [y * col_chunk, (y + 1) * col_chunk) -- the last chunk also
  // takes the remainder -- and leaves its PARTIAL result in the first d columns of its own (now dead) block of W
  // (`parts` == W); `reduce_col_parts_kernel` then adds the chunks in ascending order.  The decomposition depends on m2
  // only, never on the number of rows, so results stay bit-identical however a t-batch is chunked.
  constexpr int DP = 8 * NTD;
  extern __shared__ __align__(16) double sm[];
  double* s1 = sm;                         // [d][BW_T]      U1^T tile (rows of this CTA)
  double* s2 = s1 + (size_t)d * BW_T;      // [DP][BW_P]     [U2 | 1 | 0]^T tile
  double* sG = s2 + (size_t)DP * BW_P;     // [BW_T][BW_P]   G tile
  double* sc = sG + (size_t)BW_T * BW_P;   // [BW_T] col_vec tile
  double* sr = sc + BW_T;                  // [BW_T] row_scale
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t row0 = (int64_t)blockIdx.x * BW_T;

  for (int idx = tid; idx < BW_T * d; idx += 256) {
    int p = idx / d, k = idx - p * d;
    int64_t gr = row0 + p;
    s1[k * BW_T + p] = (gr < m1) ? U1[gr * d + k] : 0.0;
  }
  for (int p = tid; p < BW_T; p += 256) {
    int64_t gr = row0 + p;
    sr[p] = (row_scale != nullptr && gr < m1) ? row_scale[gr] : 0.0;
  }
  // constant rows of the B operand: row d = ones (row sums), rows d+1.. = zeros
  for (int idx = tid; idx < (DP - d) * BW_P; idx += 256) s2[(size_t)d * BW_P + idx] = (idx < BW_P) ? 1.0 : 0.0;

  double vacc[NTD][2];
#pragma unroll
  for (int j = 0; j < NTD; j++) { vacc[j][0] = 0.0; vacc[j][1] = 0.0; }
  const int stage_p0 = tid / d, stage_k0 = tid - stage_p0 * d, stage_dq = 256 / d, stage_dr = 256 - stage_dq * d;

  const int n_chunks = (col_chunk > 0) ? (int)gridDim.y : 1;
  const int col_begin = (col_chunk > 0) ? (int)blockIdx.y * col_chunk : 0;
  const int col_end = (col_chunk > 0 && (int)blockIdx.y + 1 < n_chunks) ? col_begin + col_chunk : m2;
  for (int c0 = col_begin; c0 < col_end; c0 += BW_T) {
    __syncthreads();  // previous tile's DMMA reads of s2 / sG are done
    {
      const double* src = U2 + (int64_t)c0 * d;   // element idx of the tile lives at c0 * d + idx (no division per element)
      int pp = stage_p0, kk = stage_k0;
      for (int idx = tid; idx < BW_T * d; idx += 256) {
        s2[kk * BW_P + pp] = (c0 + pp < col_end) ? src[idx] : 0.0;
        kk += stage_dr; pp += stage_dq;
        if (kk >= d) { kk -= d; pp++; }
      }
    }
    for (int p = tid; p < BW_T; p += 256) {
      int gc = c0 + p;
      sc[p] = (col_vec != nullptr && gc < col_end) ? col_vec[gc] : 0.0;
    }
    // the 4 x 4 block of upstream gradients W is requested before the barrier and the distance loop so that its HBM
    // latency is covered by them
    double w[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int64_t gr = row0 + ty * 4 + i;
      const int gc0 = c0 + tx * 4;
#pragma unroll
      for (int j = 0; j < 4; j++) w[i][j] = 0.0;
      if (gr < m1) {
        const double* wp = W + gr * ldw + gc0;
        if (gc0 + 3 < col_end && ((ldw & 1) == 0)) {
          double2 w01 = *reinterpret_cast<const double2*>(wp);
          double2 w23 = *reinterpret_cast<const double2*>(wp + 2);
          w[i][0] = w01.x; w[i][1] = w01.y; w[i][2] = w23.x; w[i][3] = w23.y;
        } else {
#pragma unroll
          for (int j = 0; j < 4; j++) if (gc0 + j < col_end) w[i][j] = wp[j];
        }
      }
    }
    __syncthreads();
    // ---- phase 1
    double sq[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) sq[i][j] = 0.0;
    for (int k = 0; k < d; k++) {
      const double4 a4 = *reinterpret_cast<const double4*>(s1 + k * BW_T + ty * 4);
      const double4 b4 = *reinterpret_cast<const double4*>(s2 + k * BW_P + tx * 4);
      const double a[4] = {a4.x, a4.y, a4.z, a4.w};
      const double bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
          double df = a[i] - bb[j];
          sq[i][j] = fma(df, df, sq[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int64_t gr = row0 + ty * 4 + i;
      const int gc0 = c0 + tx * 4;
      const double rs = sr[ty * 4 + i];
      double gv[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const bool ok = (gr < m1) && (gc0 + j < col_end);
        gv[j] = ok ? (w[i][j] + rs * sc[tx * 4 + j]) * kernel_dfactor(kernel_id, outputscale, sq[i][j]) : 0.0;
      }
      *reinterpret_cast<double4*>(sG + (ty * 4 + i) * BW_P + tx * 4) = make_double4(gv[0], gv[1], gv[2], gv[3]);
    }
    __syncthreads();
    // ---- phase 2: warp `warp` owns rows 8*warp .. 8*warp+7
    const double* ga = sG + (warp * 8 + g) * BW_P + t4;
    const double* ub = s2 + (size_t)g * BW_P + t4;
#pragma unroll 4
    for (int kk = 0; kk < BW_T; kk += 4) {
      const double a = ga[kk];
#pragma unroll
      for (int j = 0; j < NTD; j++) dmma884c(vacc[j][0], vacc[j][1], a, ub[(size_t)j * 8 * BW_P + kk]);
    }
  }
  __syncthreads();
  // ---- finalize: V -> shared, dU1[row][k] = u1[row][k] * rowsum[row] - V[row][k]
  double* sV = s2;  // [BW_T][DP]: reuses the (now idle) operand tile AND the G tile behind it -- DP can exceed BW_P (d = 64)
#pragma unroll
  for (int j = 0; j < NTD; j++) {
    sV[(warp * 8 + g) * DP + j * 8 + 2 * t4] = vacc[j][0];
    sV[(warp * 8 + g) * DP + j * 8 + 2 * t4 + 1] = vacc[j][1];
  }
  __syncthreads();
  for (int idx = tid; idx < BW_T * d; idx += 256) {
    int p = idx / d, k = idx - p * d;
    int64_t gr = row0 + p;
    if (gr < m1) {
      double v = s1[k * BW_T + p] * sV[p * DP + d] - sV[p * DP + k];
      if (col_chunk > 0) {
        parts[gr * ldw + col_begin + k] = v;
      } else {
        double* dst = dU1 + gr * d + k;
        *dst = accumulate ? (*dst + v) : v;
      }
    }
  }
}

}  // namespace mcacq

extern "C" int mcacq_scale_inputs(const double* X, int64_t rows, int d, const double* x_offset,
                                  const double* x_coef, const double* lengthscale, double* U, void* stream) {
  using namespace mcacq;
  if (!X || !U || !x_offset || !x_coef || !lengthscale || rows < 0 || d <= 0) return MCACQ_EINVAL;
  if (rows == 0) return 0;
  int64_t total = rows * d;
  int threads = 256;
  int64_t blocks = (total + threads - 1) / threads;
  scale_inputs_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(X, total, d, x_offset, x_coef, lengthscale, U);
  count_launch();
  MCACQ_CUDA_CHECK_LAUNCH();
  return 0;
}

namespace mcacq {
__global__ void fill_value_kernel(double* p, int64_t n, double v) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
int fill_value(double* p, int64_t n, double v, cudaStream_t st) {
  if (n == 0) return 0;
  fill_value_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, n, v);
  count_launch();
  MCACQ_CUDA_CHECK_LAUNCH();
  return 0;
}
int unscale_grad(const double* dU, int64_t rows, int d, const double* coef, const double* ls, double* dX,
                 cudaStream_t st) {
  int64_t total = rows * d;
  if (total == 0) return 0;
  int threads = 256;
  int64_t blocks = (total + threads - 1) / threads;
  unscale_grad_kernel<<<(unsigned)blocks, threads, 0, st>>>(dU, total, d, coef, ls, dX);
  count_launch();
  MCACQ_CUDA_CHECK_LAUNCH();
  return 0;
}
}  // namespace mcacq

namespace mcacq {
static int cov_cross_launch(bool sliced, int kernel_id, double outputscale, const double* U1, int64_t m1, const double* U2,
                            int m2, int d, double* K, int64_t ldk, int8_t* S, int G, int fixed_exp, const double* alpha,
                            double* mean_part, cudaStream_t st) {
  if (d > MCACQ_MAX_D) return MCACQ_ELIMIT;
  if (kernel_id != MCACQ_KERNEL_RBF && kernel_id != MCACQ_KERNEL_MATERN52) return MCACQ_EINVAL;
  if (m1 == 0 || ldk == 0) return 0;
  int64_t gy = (m1 + CT - 1) / CT;
  int gx = (int)((ldk + CT - 1) / CT);
  if (sliced) gx = (gx + CT_PER_CTA - 1) / CT_PER_CTA;
  size_t smem = (size_t)2 * d * CT * sizeof(double);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(cov_cross_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * MCACQ_MAX_D * CT * 8);
    cudaFuncSetAttribute(cov_cross_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * MCACQ_MAX_D * CT * 8);
    attr = true;
  }
  // gridDim.y is limited to 65535: split very tall problems
  const int64_t max_gy = 65535;
  for (int64_t y0 = 0; y0 < gy; y0 += max_gy) {
    int64_t ny = (gy - y0 < max_gy) ? gy - y0 : max_gy;
    int64_t r0 = y0 * CT;
    dim3 grid(gx, (unsigned)ny);
    if (sliced)
      cov_cross_kernel<true><<<grid, 256, smem, st>>>(kernel_id, outputscale, U1 + r0 * d, m1 - r0, U2, m2, d, nullptr, ldk,
                                                      S, G, fixed_exp, alpha, mean_part, m1, r0);
    else
      cov_cross_kernel<false><<<grid, 256, smem, st>>>(kernel_id, outputscale, U1 + r0 * d, m1 - r0, U2, m2, d,
                                                       K + r0 * ldk, ldk, nullptr, 0, 0, nullptr, nullptr, m1, r0);
    count_launch();
  }
  MCACQ_CUDA_CHECK_LAUNCH();
  return 0;
}
}  // namespace mcacq

extern "C" int mcacq_cov_cross(int kernel_id, double outputscale, const double* U1, int64_t m1, const double* U2,
                               int m2, int d, double* K, int64_t ldk, void* stream) {
  if (!U1 || !U2 || !K || m1 < 0 || m2 < 0 || d <= 0 || ldk < m2) return MCACQ_EINVAL;
  return mcacq::cov_cross_launch(false, kernel_id, outputscale, U1, m1, U2, m2, d, K, ldk, nullptr, 0, 0, nullptr, nullptr,
                                 (cudaStream_t)stream);
}

extern "C" int mcacq_cov_cross_sliced(int kernel_id, double outputscale, const double* U1, int64_t m1, const double* U2,
                                      int m2, int d, int64_t ldk, const double* alpha, int G, int fixed_exp,
                                      int8_t* slices, double* mean_part, void* stream) {
  if (!U1 || !U2 || !alpha || !slices || !mean_part || m1 < 0 || m2 < 0 || d <= 0 || ldk < m2 || (ldk % 16) != 0)
    return MCACQ_EINVAL;
  if (G < 1 || G > 7) return MCACQ_EINVAL;
  return mcacq::cov_cross_launch(true, kernel_id, outputscale, U1, m1, U2, m2, d, nullptr, ldk, slices, G, fixed_exp, alpha,
                                 mean_part, (cudaStream_t)stream);
}

namespace mcacq {
__global__ void reduce_col_parts_kernel(const double* __restrict__ parts, int64_t ldw, int n_chunks, int col_chunk,
                                        int64_t m1, int d, double* __restrict__ dU1, int accumulate) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= m1 * d) return;
  const int64_t row = idx / d;
  const int k = (int)(idx - row * d);
  double v = accumulate ? dU1[idx] : 0.0;
  for (int c = 0; c < n_chunks; c++) v += parts[row * ldw + (int64_t)c * col_chunk + k];
  dU1[idx] = v;
}

template <int NTD>
static int launch_cov_bwd(int kernel_id, double outputscale, const double* U1, int64_t m1, const double* U2, int m2,
                          int d, const double* W, int64_t ldw, const double* row_scale, const double* col_vec,
                          double* dU1, int accumulate, double* parts, cudaStream_t st) {
  constexpr int DP = 8 * NTD;
  size_t smem = ((size_t)d * BW_T + (size_t)DP * BW_P + (size_t)BW_T * BW_P + 2 * BW_T) * sizeof(double);
  auto kern = cov_cross_bwd_kernel<NTD>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int64_t blocks = (m1 + BW_T - 1) / BW_T;
  const int n_chunks = (parts != nullptr) ? m2 / BW_CHUNK : 1;   // every chunk >= BW_CHUNK >= d columns wide
  if (n_chunks >= 2) {
    kern<<<dim3((unsigned)blocks, (unsigned)n_chunks), 256, smem, st>>>(kernel_id, outputscale, U1, m1, U2, m2, d, W, ldw,
                                                                         row_scale, col_vec, dU1, accumulate, BW_CHUNK, parts);
    count_launch();
    MCACQ_CUDA_CHECK_LAUNCH();
    const int64_t total = m1 * d;
    reduce_col_parts_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(parts, ldw, n_chunks, BW_CHUNK, m1, d, dU1,
                                                                              accumulate);
    count_launch();
    MCACQ_CUDA_CHECK_LAUNCH();
    return 0;
  }
  kern<<<(unsigned)blocks, 256, smem, st>>>(kernel_id, outputscale, U1, m1, U2, m2, d, W, ldw, row_scale, col_vec, dU1,
                                            accumulate, 0, nullptr);
  count_launch();
  MCACQ_CUDA_CHECK_LAUNCH();
  return 0;
}

static int cov_cross_bwd_dispatch(int kernel_id, double outputscale, const double* U1, int64_t m1, const double* U2, int m2,
                                  int d, const double* W, int64_t ldw, const double* row_scale, const double* col_vec,
                                  double* dU1, int accumulate, double* parts, cudaStream_t st) {
#define MCACQ_BWD_CASE(NT) \
  return launch_cov_bwd<NT>(kernel_id, outputscale, U1, m1, U2, m2, d, W, ldw, row_scale, col_vec, dU1, accumulate, parts, st)
  const int ntd = (d + 1 + 7) / 8;  // 8 * NTD >= d + 1 (one extra all-ones column for the row sums)
  if (ntd <= 1) MCACQ_BWD_CASE(1);
  if (ntd <= 2) MCACQ_BWD_CASE(2);
  if (ntd <= 3) MCACQ_BWD_CASE(3);
  if (ntd <= 4) MCACQ_BWD_CASE(4);
  if (ntd <= 6) MCACQ_BWD_CASE(6);
  MCACQ_BWD_CASE(9);
#undef MCACQ_BWD_CASE
}

// Fused-path variant: W is workspace that dies with this call, so the column chunks may park their partial results in it
// (see cov_cross_bwd_kernel); many more CTAs than row tiles, which is what an L-BFGS round (a few hundred rows) needs.
int cov_cross_bwd_split(int kernel_id, double outputscale, const double* U1, int64_t m1, const double* U2, int m2, int d,
                        double* W, int64_t ldw, const double* row_scale, const double* col_vec, double* dU1, int accumulate,
                        cudaStream_t st) {
  if (m1 == 0) return 0;
  return cov_cross_bwd_dispatch(kernel_id, outputscale, U1, m1, U2, m2, d, W, ldw, row_scale, col_vec, dU1, accumulate, W, st);
}
}  // namespace mcacq

extern "C" int mcacq_cov_cross_bwd(int kernel_id, double outputscale, const double* U1, int64_t m1, const double* U2,
                                   int m2, int d, const double* W, int64_t ldw, const double* row_scale,
                                   const double* col_vec, double* dU1, int accumulate, void* stream) {
  using namespace mcacq;
  if (!U1 || !U2 || !W || !dU1 || m1 < 0 || m2 < 0 || d <= 0 || ldw < m2) return MCACQ_EINVAL;
  if (d > MCACQ_MAX_D) return MCACQ_ELIMIT;
  if (m1 == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  return cov_cross_bwd_dispatch(kernel_id, outputscale, U1, m1, U2, m2, d, W, ldw, row_scale, col_vec, dU1, accumulate,
                                /*parts=*/nullptr, st);
}


namespace mcacq {
__global__ void sobol_draw_kernel(const int64_t* __restrict__ sobolstate, const int64_t* __restrict__ shift, int dim, int64_t n,
                                  int64_t first_index, double* __restrict__ out) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= n * dim) return;
  const int64_t k = idx / dim;
  const int j = (int)(idx - k * dim);
  const uint64_t i = (uint64_t)(first_index + k);
  uint64_t gray = i ^ (i >> 1);
  int64_t acc = shift[j];
  const int64_t* row = sobolstate + (int64_t)j * 30;
  while (gray != 0) {
    const int b = __ffsll((long long)gray) - 1;
    if (b < 30) acc ^= row[b];
    gray &= gray - 1;
  }
  out[idx] = (double)acc * 9.31322574615478515625e-10;  // 2^-30, exact
}
}  // namespace mcacq

extern "C" int mcacq_sobol_draw(const int64_t* sobolstate, const int64_t* shift, int dim, int64_t n, int64_t first_index,
                                double* out, void* stream) {
  using namespace mcacq;
  if (!sobolstate || !shift || !out || dim <= 0 || n < 0 || first_index < 0) return MCACQ_EINVAL;
  if (first_index + n > ((int64_t)1 << 30)) return MCACQ_ELIMIT;  // the engine has 30 bits
  if (n == 0) return 0;
  const int64_t total = n * dim;
  sobol_draw_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(sobolstate, shift, dim, n, first_index, out);
  count_launch();
  MCACQ_CUDA_CHECK_LAUNCH();
  return 0;
}
