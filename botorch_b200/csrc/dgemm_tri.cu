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
//   * CTA tile 64 x 64 x 16, 4 warps as 2(M) x 2(N), warp tile 32 x 32 = 4 x 4 DMMA tiles (32 accumulator
//     doubles per thread, <= 128 registers) so that FOUR CTAs are resident per SM: their barriers and
//     pipeline prologues interleave, which measured 32.8 TF/s against 28.6 TF/s for one 128 x 128 CTA per
//     SM (profiles/r01_gemm_variants.txt).
//   * 3-stage cp.async (LDGSTS) pipeline, padded shared-memory rows (A: 20, B: BN+4 doubles) so that
//     the 8x4 / 4x8 fragment reads of a half-warp touch all 32 banks exactly once.
//   * persistent grid (4 CTAs per SM), dynamic tile scheduler over an L2-friendly order:
//     groups of 2048 rows sweep the N-tiles from the longest k-range to the shortest, so the A
//     panels of a group (64 MB at np=4096) and the live B panels stay L2 resident.
#include "dgemm_tri.cuh"

namespace mcacq {
__global__ void zero_counter_kernel(int* c) { *c = 0; }
}  // namespace mcacq

extern "C" int mcacq_dgemm_tri(int tri_mode, int64_t M, int np, const double* A, const double* B, double* C,
                               int32_t* tile_counter, void* stream) {
  using namespace mcacq;
  if (!A || !B || !C || !tile_counter || M < 0 || np <= 0 || (np % 16) != 0) return MCACQ_EINVAL;
  if (tri_mode < 0 || tri_mode > 2) return MCACQ_EINVAL;
  if (M == 0) return 0;
  int rc = launch_dgemm_tri<64, 64, 2, 2, 3, 4, 16>(tri_mode, M, np, A, B, C, tile_counter, (cudaStream_t)stream);
  count_launch(2);
  return rc;
}
