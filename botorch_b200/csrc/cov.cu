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

__global__ void __launch_bounds__(256)
cov_cross_kernel(int kernel_id, double outputscale, const double* __restrict__ U1, int64_t m1,
                 const double* __restrict__ U2, int m2, int d, double* __restrict__ K, int64_t ldk) {
  extern __shared__ __align__(16) double sm[];
  double* s1 = sm;             // [d][CT]
  double* s2 = sm + d * CT;    // [d][CT]
  const int tid = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.y * CT;
  const int col0 = blockIdx.x * CT;

  for (int idx = tid; idx < CT * d; idx += 256) {
    int p = idx / d, k = idx - p * d;  // coalesced read of point-major global rows
    int64_t gr = row0 + p;
    s1[k * CT + p] = (gr < m1) ? U1[gr * d + k] : 0.0;
    int gc = col0 + p;
    s2[k * CT + p] = (gc < m2) ? U2[(int64_t)gc * d + k] : 0.0;
  }
  __syncthreads();

  const int tx = tid & 15, ty = tid >> 4;
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
    if (gr >= m1) continue;
    double v[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int gc = col0 + tx * 4 + j;
      v[j] = (gc < m2) ? kernel_value(kernel_id, outputscale, sq[i][j]) : 0.0;
    }
    int gc0 = col0 + tx * 4;
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

// ---------------------------------------------------------------------------------------------
// Backward: dU1[i][:] (+)= sum_j (W[i][j] + row_scale[i] * col_vec[j]) * g(sq_ij) * (u1_i - u2_j).
// One warp owns RPW rows; lanes stride over the columns of a 128-wide U2 tile staged dimension-major
// in shared memory; per-dimension partial sums live in registers and are shuffle-reduced at the end.
constexpr int BW_COLS = 128;
constexpr int BW_WARPS = 8;

template <int DMAX, bool KEEP_DF>
__global__ void __launch_bounds__(BW_WARPS * 32)
cov_cross_bwd_kernel(int kernel_id, double outputscale, const double* __restrict__ U1, int64_t m1,
                     const double* __restrict__ U2, int m2, int d, const double* __restrict__ W, int64_t ldw,
                     const double* __restrict__ row_scale, const double* __restrict__ col_vec,
                     double* __restrict__ dU1, int accumulate) {
  extern __shared__ __align__(16) double sm[];
  double* s2 = sm;                        // [d][BW_COLS]
  double* sc = sm + (size_t)d * BW_COLS;  // [BW_COLS] col_vec tile
  double* su = sc + BW_COLS;              // [BW_WARPS][d] row of U1 per warp (only when !KEEP_DF)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t gr = (int64_t)blockIdx.x * BW_WARPS + warp;
  const bool row_ok = gr < m1;

  double u1[KEEP_DF ? DMAX : 1];
  double acc[DMAX];
  const double rs = (row_scale != nullptr && row_ok) ? row_scale[gr] : 0.0;
#pragma unroll
  for (int k = 0; k < DMAX; k++) {
    acc[k] = 0.0;
    if (KEEP_DF) u1[k] = (k < d && row_ok) ? U1[gr * d + k] : 0.0;
  }
  if (!KEEP_DF) {
    for (int k = lane; k < d; k += 32) su[warp * d + k] = row_ok ? U1[gr * d + k] : 0.0;
  }

  for (int c0 = 0; c0 < m2; c0 += BW_COLS) {
    __syncthreads();
    for (int idx = tid; idx < BW_COLS * d; idx += BW_WARPS * 32) {
      int p = idx / d, k = idx - p * d;
      int gc = c0 + p;
      s2[k * BW_COLS + p] = (gc < m2) ? U2[(int64_t)gc * d + k] : 0.0;
    }
    for (int p = tid; p < BW_COLS; p += BW_WARPS * 32) {
      int gc = c0 + p;
      sc[p] = (col_vec != nullptr && gc < m2) ? col_vec[gc] : 0.0;
    }
    __syncthreads();
    if (!row_ok) continue;
#pragma unroll
    for (int cc = 0; cc < BW_COLS; cc += 32) {
      const int p = cc + lane;
      const int gc = c0 + p;
      if (gc < m2) {
        const double w = W[gr * ldw + gc] + rs * sc[p];
        double sq = 0.0;
        if (KEEP_DF) {
          double df[DMAX];
#pragma unroll
          for (int k = 0; k < DMAX; k++) {
            if (k < d) {
              df[k] = u1[k] - s2[k * BW_COLS + p];
              sq = fma(df[k], df[k], sq);
            }
          }
          const double wg = w * kernel_dfactor(kernel_id, outputscale, sq);
#pragma unroll
          for (int k = 0; k < DMAX; k++)
            if (k < d) acc[k] = fma(wg, df[k], acc[k]);
        } else {
#pragma unroll
          for (int k = 0; k < DMAX; k++) {
            if (k < d) {
              double df = su[warp * d + k] - s2[k * BW_COLS + p];
              sq = fma(df, df, sq);
            }
          }
          const double wg = w * kernel_dfactor(kernel_id, outputscale, sq);
#pragma unroll
          for (int k = 0; k < DMAX; k++)
            if (k < d) acc[k] = fma(wg, su[warp * d + k] - s2[k * BW_COLS + p], acc[k]);
        }
      }
    }
  }

#pragma unroll
  for (int k = 0; k < DMAX; k++) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && k < d && row_ok) {
      double* dst = dU1 + gr * d + k;
      *dst = accumulate ? (*dst + v) : v;
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

extern "C" int mcacq_cov_cross(int kernel_id, double outputscale, const double* U1, int64_t m1, const double* U2,
                               int m2, int d, double* K, int64_t ldk, void* stream) {
  using namespace mcacq;
  if (!U1 || !U2 || !K || m1 < 0 || m2 < 0 || d <= 0 || ldk < m2) return MCACQ_EINVAL;
  if (d > MCACQ_MAX_D) return MCACQ_ELIMIT;
  if (kernel_id != MCACQ_KERNEL_RBF && kernel_id != MCACQ_KERNEL_MATERN52) return MCACQ_EINVAL;
  if (m1 == 0 || ldk == 0) return 0;
  int64_t gy = (m1 + CT - 1) / CT;
  int gx = (int)((ldk + CT - 1) / CT);
  size_t smem = (size_t)2 * d * CT * sizeof(double);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(cov_cross_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * MCACQ_MAX_D * CT * 8);
    attr = true;
  }
  // gridDim.y is limited to 65535: split very tall problems
  const int64_t max_gy = 65535;
  for (int64_t y0 = 0; y0 < gy; y0 += max_gy) {
    int64_t ny = (gy - y0 < max_gy) ? gy - y0 : max_gy;
    int64_t r0 = y0 * CT;
    dim3 grid(gx, (unsigned)ny);
    cov_cross_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(kernel_id, outputscale, U1 + r0 * d, m1 - r0, U2, m2, d,
                                                                 K + r0 * ldk, ldk);
    count_launch();
  }
  MCACQ_CUDA_CHECK_LAUNCH();
  return 0;
}

namespace mcacq {
template <int DMAX, bool KEEP_DF>
static int launch_cov_bwd(int kernel_id, double outputscale, const double* U1, int64_t m1, const double* U2, int m2,
                          int d, const double* W, int64_t ldw, const double* row_scale, const double* col_vec,
                          double* dU1, int accumulate, cudaStream_t st) {
  size_t smem = ((size_t)d * BW_COLS + BW_COLS + (size_t)BW_WARPS * d) * sizeof(double);
  auto kern = cov_cross_bwd_kernel<DMAX, KEEP_DF>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int64_t rows_per_cta = (int64_t)BW_WARPS;
  int64_t blocks = (m1 + rows_per_cta - 1) / rows_per_cta;
  kern<<<(unsigned)blocks, BW_WARPS * 32, smem, st>>>(kernel_id, outputscale, U1, m1, U2, m2, d, W, ldw, row_scale,
                                                      col_vec, dU1, accumulate);
  count_launch();
  MCACQ_CUDA_CHECK_LAUNCH();
  return 0;
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
#define MCACQ_BWD_CASE(DM, RP) \
  return launch_cov_bwd<DM, RP>(kernel_id, outputscale, U1, m1, U2, m2, d, W, ldw, row_scale, col_vec, dU1, accumulate, st)
  if (d <= 8) MCACQ_BWD_CASE(8, true);
  if (d <= 16) MCACQ_BWD_CASE(16, true);
  if (d <= 24) MCACQ_BWD_CASE(24, true);
  if (d <= 32) MCACQ_BWD_CASE(32, true);
  if (d <= 48) MCACQ_BWD_CASE(48, false);
  MCACQ_BWD_CASE(64, false);
#undef MCACQ_BWD_CASE
}
