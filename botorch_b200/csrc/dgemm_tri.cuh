// Blocked FP64 tensor-core contraction against the cached triangular factor.
//
//   C[M x np] = A[M x np] * B[np x np]       (row-major, pitch np, np % 16 == 0)
//
// B is the dense-stored triangular `covar_cache` R = L^{-T} (upper) on the forward pass
// (A_out = K(X, X_train) R, the reference's `test_train_covar @ covar_cache`, gpytorch
// exact_predictive_covar under fast_pred_var) or its transpose (lower) on the backward pass
// (dK = dA R^T).  Only the k-range that meets non-zeros of B is contracted, so the kernel executes
// M*np*(np+128) FMAs instead of the reference's dense 2*M*np^2/2.
//
// Hardware mapping and tile-shape rationale: see dgemm_tri.cu.
#pragma once
#include "common.cuh"

namespace mcacq {

constexpr int GROUP_M_ROWS = 2048;  // rows per L2 super-row (16 tiles of 128)

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int sz = pred ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int BM, int BN, int WARPS_M, int WARPS_N, int STAGES, int BK>
struct GemmCfg {
  static constexpr int A_LD = BK + 4;  // 160 B / 288 B row stride: conflict-free 8x4 fragment reads
  static constexpr int THREADS = WARPS_M * WARPS_N * 32;
  static constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;
  static constexpr int MT = WM / 8, NT = WN / 8;
  static constexpr int B_LD = BN + 4;
  static constexpr int A_STAGE = BM * A_LD, B_STAGE = BK * B_LD;
  static constexpr size_t SMEM = (size_t)STAGES * (A_STAGE + B_STAGE) * sizeof(double);
  static constexpr int GROUP_M = GROUP_M_ROWS / BM;
};

template <int BM, int BN, int WARPS_M, int WARPS_N, int STAGES, int MINB, int BK = 16>
__global__ void __launch_bounds__(WARPS_M * WARPS_N * 32, MINB)
dgemm_tri_kernel(int tri_mode, int64_t M, int np, const double* __restrict__ A, const double* __restrict__ B,
                 double* __restrict__ C, int* __restrict__ tile_counter) {
  using Cfg = GemmCfg<BM, BN, WARPS_M, WARPS_N, STAGES, BK>;
  constexpr int THREADS = Cfg::THREADS, MT = Cfg::MT, NT = Cfg::NT, B_LD = Cfg::B_LD, A_LD = Cfg::A_LD;
  constexpr int A_STAGE = Cfg::A_STAGE, B_STAGE = Cfg::B_STAGE, GROUP_M = Cfg::GROUP_M;
  extern __shared__ __align__(16) double smem[];
  double* sA = smem;
  double* sB = smem + STAGES * A_STAGE;
  __shared__ int s_tile;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int wm = warp / WARPS_N;
  const int wn = warp % WARPS_N;
  const int g = lane >> 2;   // 0..7
  const int t4 = lane & 3;   // 0..3

  const int64_t m_tiles = (M + BM - 1) / BM;
  const int n_tiles = (np + BN - 1) / BN;
  const int k_tiles_total = np / BK;
  const int64_t total_tiles = m_tiles * n_tiles;

  for (;;) {
    __syncthreads();  // previous tile's smem reads are done; s_tile reuse
    if (tid == 0) s_tile = atomicAdd(tile_counter, 1);
    __syncthreads();
    const int64_t tile = s_tile;
    if (tile >= total_tiles) break;

    // ---- tile order: (M-group, N-tile by descending k-range, M-tile within the group)
    const int64_t group_sz = (int64_t)GROUP_M * n_tiles;
    const int64_t grp = tile / group_sz;
    const int64_t m0_tile = grp * GROUP_M;
    const int64_t rows_in_group = (m_tiles - m0_tile < GROUP_M) ? (m_tiles - m0_tile) : GROUP_M;
    const int64_t within = tile - grp * group_sz;
    const int nrank = (int)(within / rows_in_group);
    const int64_t mt = m0_tile + within % rows_in_group;
    int nt;
    if (tri_mode == MCACQ_TRI_LOWER) nt = nrank;            // smallest column tile has the longest k-range
    else nt = n_tiles - 1 - nrank;                          // upper / dense: largest column tile first

    int kt_begin = 0, kt_end = k_tiles_total;
    if (tri_mode == MCACQ_TRI_UPPER) {
      int kmax = (nt + 1) * BN; if (kmax > np) kmax = np;
      kt_end = kmax / BK;
    } else if (tri_mode == MCACQ_TRI_LOWER) {
      kt_begin = (nt * BN) / BK;
    }

    const int64_t row0 = mt * BM;
    const int col0 = nt * BN;

    double acc[MT][NT][2];
#pragma unroll
    for (int i = 0; i < MT; i++)
#pragma unroll
      for (int j = 0; j < NT; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    // ---- async tile loader: A BM x 16 (BM*8 16-byte chunks), B 16 x BN (8*BN chunks)
    auto load_stage = [&](int stage, int kt) {
      double* a_dst = sA + stage * A_STAGE;
      double* b_dst = sB + stage * B_STAGE;
      const int k0 = kt * BK;
#pragma unroll
      for (int c = tid; c < BM * (BK / 2); c += THREADS) {
        int r = c / (BK / 2), ch = c % (BK / 2);
        int64_t gr = row0 + r;
        bool ok = gr < M;
        const double* src = A + (ok ? gr : 0) * (int64_t)np + k0 + ch * 2;
        cp_async16(a_dst + r * A_LD + ch * 2, src, ok);
      }
#pragma unroll
      for (int c = tid; c < BK * (BN / 2); c += THREADS) {
        int r = c / (BN / 2), ch = c % (BN / 2);
        int gc = col0 + ch * 2;
        bool ok = gc < np;
        const double* src = B + (int64_t)(k0 + r) * np + (ok ? gc : 0);
        cp_async16(b_dst + r * B_LD + ch * 2, src, ok);
      }
    };

    const int nk = kt_end - kt_begin;
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
      if (s < nk) load_stage(s, kt_begin + s);
      cp_async_commit();
    }

    for (int it = 0; it < nk; it++) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      {
        int nxt = it + STAGES - 1;
        if (nxt < nk) load_stage(nxt % STAGES, kt_begin + nxt);
        cp_async_commit();
      }
      const double* a_s = sA + (it % STAGES) * A_STAGE + (wm * Cfg::WM + g) * A_LD + t4;
      const double* b_s = sB + (it % STAGES) * B_STAGE + t4 * B_LD + wn * Cfg::WN + g;
#pragma unroll
      for (int kk = 0; kk < BK; kk += 4) {
        double af[MT], bf[NT];
#pragma unroll
        for (int i = 0; i < MT; i++) af[i] = a_s[i * 8 * A_LD + kk];
#pragma unroll
        for (int j = 0; j < NT; j++) bf[j] = b_s[kk * B_LD + j * 8];
#pragma unroll
        for (int i = 0; i < MT; i++)
#pragma unroll
          for (int j = 0; j < NT; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      }
    }
    cp_async_wait<0>();

    // ---- epilogue: each thread owns C[row g][cols 2*t4, 2*t4+1] of every 8x8 tile (16B stores)
#pragma unroll
    for (int i = 0; i < MT; i++) {
      int64_t gr = row0 + wm * Cfg::WM + i * 8 + g;
      if (gr < M) {
#pragma unroll
        for (int j = 0; j < NT; j++) {
          int gc = col0 + wn * Cfg::WN + j * 8 + t4 * 2;
          if (gc < np) {
            double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
            *reinterpret_cast<double2*>(C + gr * (int64_t)np + gc) = v;
          }
        }
      }
    }
  }
}

__global__ void zero_counter_kernel(int* c);

template <int BM, int BN, int WARPS_M, int WARPS_N, int STAGES, int MINB, int BK = 16>
int launch_dgemm_tri(int tri_mode, int64_t M, int np, const double* A, const double* B, double* C,
                     int32_t* tile_counter, cudaStream_t st) {
  using Cfg = GemmCfg<BM, BN, WARPS_M, WARPS_N, STAGES, BK>;
  auto kern = dgemm_tri_kernel<BM, BN, WARPS_M, WARPS_N, STAGES, MINB, BK>;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    if (e != cudaSuccess) { sms = 0; return (int)e; }
  }
  int64_t m_tiles = (M + BM - 1) / BM;
  int n_tiles = (np + BN - 1) / BN;
  int64_t total = m_tiles * n_tiles;
  int grid = (int)((total < (int64_t)sms * MINB) ? total : (int64_t)sms * MINB);
  zero_counter_kernel<<<1, 1, 0, st>>>(tile_counter);
  kern<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(tri_mode, M, np, A, B, C, tile_counter);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace mcacq
