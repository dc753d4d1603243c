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
// Hardware mapping (B200, sm_100a): there is no tcgen05 FP64 kind -- FP64 tensor math is the
// warp-level `mma.sync.m8n8k4.f64` (SASS DMMA.8x8x4; m16n8k{4,8,16} are decomposed by ptxas into
// the same instruction).  Measured pipe peak 37.2 TF/s (profiles/r01_ubench_fp64_pipe.txt).
//   * CTA tile 128 x 128 x 16, 8 warps as 2(M) x 4(N), warp tile 64 x 32 = 8 x 4 DMMA tiles,
//     64 accumulator doubles per thread.
//   * 4-stage cp.async (LDGSTS) pipeline, padded shared-memory rows (A: 20, B: 132 doubles) so that
//     the 8x4 / 4x8 fragment reads of a half-warp touch all 32 banks exactly once.
//   * persistent grid (one CTA per SM), dynamic tile scheduler over an L2-friendly order:
//     groups of 16 M-tiles sweep the N-tiles from the longest k-range to the shortest, so the A
//     panels of a group (64 MB at np=4096) and the live B panels stay L2 resident.
#include "common.cuh"

namespace mcacq {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int STAGES = 4;
constexpr int A_LD = BK + 4;    // 20 doubles = 160 B row stride
constexpr int B_LD = BN + 4;    // 132 doubles = 1056 B row stride
constexpr int A_STAGE = BM * A_LD;  // doubles
constexpr int B_STAGE = BK * B_LD;
constexpr int GEMM_THREADS = 256;
constexpr int GROUP_M = 16;
constexpr size_t GEMM_SMEM = (size_t)STAGES * (A_STAGE + B_STAGE) * sizeof(double) + 16;

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

__global__ void __launch_bounds__(GEMM_THREADS, 1)
dgemm_tri_kernel(int tri_mode, int64_t M, int np, const double* __restrict__ A, const double* __restrict__ B,
                 double* __restrict__ C, int* __restrict__ tile_counter) {
  extern __shared__ __align__(16) double smem[];
  double* sA = smem;
  double* sB = smem + STAGES * A_STAGE;
  __shared__ int s_tile;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int wm = warp >> 2;  // 0..1  -> rows  wm*64
  const int wn = warp & 3;   // 0..3  -> cols  wn*32
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
    // the last group may be short: re-derive with its own row count
    const int nrank = (int)(within / rows_in_group);
    const int64_t mt = m0_tile + within % rows_in_group;
    if (nrank >= n_tiles) continue;  // padding slots of a short last group
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

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    // ---- async tile loader: A 128x16 (1024 x 16B chunks), B 16x128 (1024 chunks); 4 + 4 per thread
    auto load_stage = [&](int stage, int kt) {
      double* a_dst = sA + stage * A_STAGE;
      double* b_dst = sB + stage * B_STAGE;
      const int k0 = kt * BK;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        int c = tid + i * GEMM_THREADS;  // 0..1023
        int r = c >> 3, ch = c & 7;      // row 0..127, 16B chunk 0..7
        int64_t gr = row0 + r;
        bool ok = gr < M;
        const double* src = A + (ok ? gr : 0) * (int64_t)np + k0 + ch * 2;
        cp_async16(a_dst + r * A_LD + ch * 2, src, ok);
      }
#pragma unroll
      for (int i = 0; i < 4; i++) {
        int c = tid + i * GEMM_THREADS;
        int r = c >> 6, ch = c & 63;  // k row 0..15, 16B chunk 0..63
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
      const double* a_s = sA + (it % STAGES) * A_STAGE + (wm * 64 + g) * A_LD + t4;
      const double* b_s = sB + (it % STAGES) * B_STAGE + t4 * B_LD + wn * 32 + g;
#pragma unroll
      for (int kk = 0; kk < BK; kk += 4) {
        double af[8], bf[4];
#pragma unroll
        for (int i = 0; i < 8; i++) af[i] = a_s[i * 8 * A_LD + kk];
#pragma unroll
        for (int j = 0; j < 4; j++) bf[j] = b_s[kk * B_LD + j * 8];
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      }
    }
    cp_async_wait<0>();

    // ---- epilogue: each thread owns C[row g][cols 2*t4, 2*t4+1] of every 8x8 tile (16B stores)
#pragma unroll
    for (int i = 0; i < 8; i++) {
      int64_t gr = row0 + wm * 64 + i * 8 + g;
      if (gr < M) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
          int gc = col0 + wn * 32 + j * 8 + t4 * 2;
          if (gc < np) {
            double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
            *reinterpret_cast<double2*>(C + gr * (int64_t)np + gc) = v;
          }
        }
      }
    }
  }
}

__global__ void zero_counter_kernel(int* c) { *c = 0; }

}  // namespace mcacq

extern "C" int mcacq_dgemm_tri(int tri_mode, int64_t M, int np, const double* A, const double* B, double* C,
                               int32_t* tile_counter, void* stream) {
  using namespace mcacq;
  if (!A || !B || !C || !tile_counter || M < 0 || np <= 0 || (np % 16) != 0) return MCACQ_EINVAL;
  if (tri_mode < 0 || tri_mode > 2) return MCACQ_EINVAL;
  if (M == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  static int sms = 0;
  static bool attr_set = false;
  if (!attr_set) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaFuncSetAttribute(dgemm_tri_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  int64_t m_tiles = (M + BM - 1) / BM;
  int n_tiles = (np + BN - 1) / BN;
  int grid = (int)((m_tiles * n_tiles < sms) ? m_tiles * n_tiles : sms);
  zero_counter_kernel<<<1, 1, 0, st>>>(tile_counter);
  dgemm_tri_kernel<<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(tri_mode, M, np, A, B, C, tile_counter);
  count_launch(2);
  MCACQ_CUDA_CHECK_LAUNCH();
  return 0;
}
