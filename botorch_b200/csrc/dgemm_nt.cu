// General FP64 DMMA contraction in "NT" form:  C[M x N] = A[M x K] * Bt[N x K]^T   (both operands K-contiguous).
//
// Used by the joint-posterior route (MaxPosteriorSampling / Thompson sampling over N
// candidates, reference botorch/generation/sampling.py:89-155):
//   * SYRK   Sigma -= A A^T          (A = K(X, X_train) R, N x np):       Bt = A,  a_lower = 0
//   * TRMM   Y = L Z^T               (L = chol(Sigma) lower, Z = base samples [num_samples x N]):  a_lower = 1
//     (only k <= row-tile end is contracted, halving the work).
// Same tile machinery as dgemm_tri.cu: 64 x 64 x 16 CTA tiles, 4 warps of 32 x 32, DMMA.8x8x4, 3-stage cp.async,
// 4 CTAs per SM, persistent grid with an atomic tile counter.  Both shared-memory tiles use the k-padded row
// layout (pitch 20 doubles), so A- and B-fragment reads are conflict-free.
#include "dgemm_tri.cuh"

namespace mcacq {

constexpr int NT_BM = 64, NT_BN = 64, NT_BK = 16, NT_STAGES = 3, NT_LD = NT_BK + 4;
constexpr int NT_THREADS = 128;
constexpr size_t NT_SMEM = (size_t)NT_STAGES * (NT_BM + NT_BN) * NT_LD * sizeof(double);

__global__ void __launch_bounds__(NT_THREADS, 4)
dgemm_nt_kernel(int a_lower, int64_t M, int N, int K, const double* __restrict__ A, int64_t lda,
                const double* __restrict__ Bt, int64_t ldb, double* C, int64_t ldc,
                int* __restrict__ tile_counter, double syrk_scale) {
  extern __shared__ __align__(16) double smem[];
  double* sA = smem;
  double* sB = smem + NT_STAGES * NT_BM * NT_LD;
  __shared__ int s_tile;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp >> 1, wn = warp & 1;
  const int g = lane >> 2, t4 = lane & 3;
  const int64_t m_tiles = (M + NT_BM - 1) / NT_BM;
  const int n_tiles = (N + NT_BN - 1) / NT_BN;
  const int k_tiles_total = (K + NT_BK - 1) / NT_BK;
  // a_lower == 2: symmetric update C := syrk_scale * (C - A A^T) (Bt == A, M == N, C symmetric on entry): only the tiles on
  // and below the diagonal are contracted, every result is stored at (row, col) and (col, row)
  const bool syrk = (a_lower == 2);
  const int64_t total = syrk ? m_tiles * (m_tiles + 1) / 2 : m_tiles * n_tiles;

  for (;;) {
    __syncthreads();
    if (tid == 0) s_tile = atomicAdd(tile_counter, 1);
    __syncthreads();
    const int64_t tile = s_tile;
    if (tile >= total) break;
    // heaviest row tiles first when A is lower triangular
    int64_t mt = (a_lower == 1) ? (m_tiles - 1 - tile / n_tiles) : tile / n_tiles;
    int nt = (int)(tile % n_tiles);
    if (syrk) {   // tile -> (mt, nt <= mt), row-major enumeration of the lower triangle
      mt = (int64_t)((sqrt(8.0 * (double)tile + 1.0) - 1.0) * 0.5);
      while (mt * (mt + 1) / 2 > tile) mt--;
      while ((mt + 1) * (mt + 2) / 2 <= tile) mt++;
      nt = (int)(tile - mt * (mt + 1) / 2);
    }
    const int64_t row0 = mt * NT_BM;
    const int col0 = nt * NT_BN;
    int kt_end = k_tiles_total;
    if (a_lower == 1) {
      int64_t kmax = row0 + NT_BM;
      if (kmax > K) kmax = K;
      kt_end = (int)((kmax + NT_BK - 1) / NT_BK);
    }
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    auto load_stage = [&](int stage, int kt) {
      double* a_dst = sA + stage * NT_BM * NT_LD;
      double* b_dst = sB + stage * NT_BN * NT_LD;
      const int k0 = kt * NT_BK;
#pragma unroll
      for (int c = tid; c < NT_BM * (NT_BK / 2); c += NT_THREADS) {
        int r = c / (NT_BK / 2), ch = c % (NT_BK / 2);
        int64_t gr = row0 + r;
        bool ok = gr < M && (k0 + ch * 2) < K;
        cp_async16(a_dst + r * NT_LD + ch * 2, A + (ok ? gr * lda + k0 + ch * 2 : 0), ok);
      }
#pragma unroll
      for (int c = tid; c < NT_BN * (NT_BK / 2); c += NT_THREADS) {
        int r = c / (NT_BK / 2), ch = c % (NT_BK / 2);
        int64_t gr = col0 + r;
        bool ok = gr < N && (k0 + ch * 2) < K;
        cp_async16(b_dst + r * NT_LD + ch * 2, Bt + (ok ? gr * ldb + k0 + ch * 2 : 0), ok);
      }
    };

    const int nk = kt_end;
#pragma unroll
    for (int s = 0; s < NT_STAGES - 1; s++) {
      if (s < nk) load_stage(s, s);
      cp_async_commit();
    }
    for (int it = 0; it < nk; it++) {
      cp_async_wait<NT_STAGES - 2>();
      __syncthreads();
      {
        int nxt = it + NT_STAGES - 1;
        if (nxt < nk) load_stage(nxt % NT_STAGES, nxt);
        cp_async_commit();
      }
      const double* a_s = sA + (it % NT_STAGES) * NT_BM * NT_LD + (wm * 32 + g) * NT_LD + t4;
      const double* b_s = sB + (it % NT_STAGES) * NT_BN * NT_LD + (wn * 32 + g) * NT_LD + t4;
#pragma unroll
      for (int kk = 0; kk < NT_BK; kk += 4) {
        double af[4], bf[4];
#pragma unroll
        for (int i = 0; i < 4; i++) af[i] = a_s[i * 8 * NT_LD + kk];
#pragma unroll
        for (int j = 0; j < 4; j++) bf[j] = b_s[j * 8 * NT_LD + kk];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int64_t gr = row0 + wm * 32 + i * 8 + g;
      if (gr < M) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
          int gc = col0 + wn * 32 + j * 8 + t4 * 2;
          double* dst = C + gr * ldc + gc;
          if (syrk) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
              if (gc + e < N) {
                const double v = syrk_scale * (dst[e] - acc[i][j][e]);
                dst[e] = v;
                if (nt != mt) C[(int64_t)(gc + e) * ldc + gr] = v;
              }
            }
            continue;
          }
          if (gc + 1 < N && ((ldc & 1) == 0)) *reinterpret_cast<double2*>(dst) = make_double2(acc[i][j][0], acc[i][j][1]);
          else {
            if (gc < N) dst[0] = acc[i][j][0];
            if (gc + 1 < N) dst[1] = acc[i][j][1];
          }
        }
      }
    }
  }
}

}  // namespace mcacq

static int dgemm_nt_launch(int a_lower, int64_t M, int N, int K, const double* A, int64_t lda, const double* Bt, int64_t ldb,
                           double* C, int64_t ldc, int32_t* tile_counter, double syrk_scale, void* stream);

namespace mcacq {
// Y[N x S] = L[N x N] Z[S x N]^T for a HANDFUL of sample vectors (S <= 8): a memory-bound sweep over the lower triangle of L.
// One warp per row, lanes stride the row (coalesced 256-byte reads of L, Z stays in L1 / L2), the S dot products share every
// L element; warp-shuffle tree at the end.  (The 64 x 64-tile DMMA kernel above spends a full tile on S = 4 columns.)
template <int SMAX>
__global__ void __launch_bounds__(256)
lower_times_few_kernel(int64_t N, int S, const double* __restrict__ L, int64_t ldl, const double* __restrict__ Z, int64_t ldz,
                       double* __restrict__ Y, int64_t ldy) {
  const int lane = threadIdx.x & 31;
  // heaviest (longest) rows first
  const int64_t row = N - 1 - ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
  if (row < 0) return;
  double acc[SMAX];
#pragma unroll
  for (int s = 0; s < SMAX; s++) acc[s] = 0.0;
  const double* lr = L + row * ldl;
  int64_t k = lane;
  for (; k + 96 <= row; k += 128) {   // four independent 256-byte row segments in flight per warp
    const double l0 = lr[k], l1 = lr[k + 32], l2 = lr[k + 64], l3 = lr[k + 96];
#pragma unroll
    for (int s = 0; s < SMAX; s++)
      if (s < S) {
        const double* z = Z + s * ldz + k;
        acc[s] = fma(l0, z[0], fma(l1, z[32], fma(l2, z[64], fma(l3, z[96], acc[s]))));
      }
  }
  for (; k <= row; k += 32) {
    const double l = lr[k];
#pragma unroll
    for (int s = 0; s < SMAX; s++)
      if (s < S) acc[s] = fma(l, Z[s * ldz + k], acc[s]);
  }
#pragma unroll
  for (int s = 0; s < SMAX; s++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[s] += __shfl_xor_sync(0xffffffffu, acc[s], o);
    if (lane == 0 && s < S) Y[row * ldy + s] = acc[s];
  }
}
}  // namespace mcacq

extern "C" int mcacq_lower_times_few(int64_t N, int S, const double* L, int64_t ldl, const double* Z, int64_t ldz, double* Y,
                                     int64_t ldy, void* stream) {
  using namespace mcacq;
  if (!L || !Z || !Y || N < 0 || S <= 0 || ldl < N || ldz < N || ldy < S) return MCACQ_EINVAL;
  if (S > 8) return MCACQ_ELIMIT;
  if (N == 0) return 0;
  const unsigned blocks = (unsigned)((N + 7) / 8);
  lower_times_few_kernel<8><<<blocks, 256, 0, (cudaStream_t)stream>>>(N, S, L, ldl, Z, ldz, Y, ldy);
  count_launch();
  MCACQ_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" int mcacq_dgemm_nt(int a_lower, int64_t M, int N, int K, const double* A, int64_t lda, const double* Bt,
                              int64_t ldb, double* C, int64_t ldc, int32_t* tile_counter, void* stream) {
  if (a_lower != 0 && a_lower != 1) return MCACQ_EINVAL;
  return dgemm_nt_launch(a_lower, M, N, K, A, lda, Bt, ldb, C, ldc, tile_counter, 1.0, stream);
}

extern "C" int mcacq_syrk_sub(int64_t N, int K, const double* A, int64_t lda, double* C, int64_t ldc, double scale,
                              int32_t* tile_counter, void* stream) {
  if (N > 0x7fffffff) return MCACQ_EINVAL;
  return dgemm_nt_launch(2, N, (int)N, K, A, lda, A, lda, C, ldc, tile_counter, scale, stream);
}

static int dgemm_nt_launch(int a_lower, int64_t M, int N, int K, const double* A, int64_t lda, const double* Bt, int64_t ldb,
                           double* C, int64_t ldc, int32_t* tile_counter, double syrk_scale, void* stream) {
  using namespace mcacq;
  if (!A || !Bt || !C || !tile_counter || M < 0 || N < 0 || K <= 0) return MCACQ_EINVAL;
  if (lda < K || ldb < K || ldc < N || (lda & 1) || (ldb & 1) || (K & 1)) return MCACQ_EINVAL;  // 16-byte cp.async chunks
  if (M == 0 || N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaFuncSetAttribute(dgemm_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NT_SMEM);
    if (e != cudaSuccess) { sms = 0; return (int)e; }
  }
  if (a_lower == 2 && (M != N || A != Bt)) return MCACQ_EINVAL;
  const int64_t mtl = (M + NT_BM - 1) / NT_BM;
  int64_t total = (a_lower == 2) ? mtl * (mtl + 1) / 2 : mtl * ((N + NT_BN - 1) / NT_BN);
  int grid = (int)((total < (int64_t)sms * 4) ? total : (int64_t)sms * 4);
  zero_counter_kernel<<<1, 1, 0, st>>>(tile_counter);
  dgemm_nt_kernel<<<grid, NT_THREADS, NT_SMEM, st>>>(a_lower, M, N, K, A, lda, Bt, ldb, C, ldc, tile_counter, syrk_scale);
  count_launch(2);
  MCACQ_CUDA_CHECK_LAUNCH();
  return 0;
}
