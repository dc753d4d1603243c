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
#include "blocks_fwd_impl.cuh"

namespace mcacq {

int posterior_blocks_fwd_q2(const BlocksParams& p, cudaStream_t st);
int posterior_blocks_fwd_q4(const BlocksParams& p, cudaStream_t st);

static int posterior_blocks_fwd_chunk(const BlocksParams& p, cudaStream_t st);

// Sxb[i][j] = s^2 (k(u_i, ub_j) - G[i][j]) in place over G = A A_base^T (same arithmetic as the tail of posterior_blocks_kernel)
__global__ void cross_finish_kernel(BlocksParams p) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t total = p.b * p.q * (int64_t)p.r;
  if (idx >= total) return;
  const int64_t i = idx / p.r;
  const int j = (int)(idx - i * p.r);
  const double* ui = p.U + i * p.d;
  const double* uj = p.U_base + (int64_t)j * p.d;
  double sq = 0.0;
  for (int k = 0; k < p.d; k++) {
    const double df = ui[k] - uj[k];
    sq = fma(df, df, sq);
  }
  p.Sxb[idx] = p.y_std * p.y_std * (kernel_value(p.kernel_id, p.outputscale, sq) - p.Sxb[idx]);
}

// Baselines beyond 64 points.  The accumulators of the cross-Gram live in registers (8 row tiles at most), and every warp
// streams the baseline rows it contracts with from L2 -- at r = 256 that is an 8 MB panel per q-batch.  So the cross term
// becomes what it is for a large baseline, a plain GEMM:  G = A A_base^T  ([b q x np] x [r x np]^T, the 64 x 64-tile DMMA
// kernel of dgemm_nt.cu, whose tiles share the baseline panel through shared memory), finished in place by
// `cross_finish_kernel`; mean / Sxx / the row maxima come from one r = 0 launch of the block kernel.  Without a scratch word
// for the GEMM's tile counter (or with MCACQ_BIGR_GEMM=0) the baseline is swept in 64-row chunks instead, one block-kernel
// launch per chunk (the first one also produces mean / Sxx, the others only their Sxb columns).
int posterior_blocks_fwd(const BlocksParams& p0, cudaStream_t st) {
  BlocksParams p = p0;
  p.r_pitch = p0.r;
  p.cross_only = 0;
  if (p0.r <= 64) return posterior_blocks_fwd_chunk(p, st);
  const char* e = getenv("MCACQ_BIGR_GEMM");
  if (p0.counter != nullptr && !(e != nullptr && atoi(e) == 0)) {
    p.r = 0; p.A_base = nullptr; p.U_base = nullptr; p.Sxb = nullptr;
    int rc = posterior_blocks_fwd_chunk(p, st);
    if (rc) return rc;
    const int64_t M = p0.b * p0.q;
    if ((rc = mcacq_dgemm_nt(0, M, p0.r, p0.np, p0.A, p0.np, p0.A_base, p0.np, p0.Sxb, p0.r, p0.counter, st))) return rc;
    const int64_t total = M * p0.r;
    cross_finish_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p0);
    count_launch();
    MCACQ_CUDA_CHECK_LAUNCH();
    return 0;
  }
  for (int r0 = 0; r0 < p0.r; r0 += 64) {
    p.r = (p0.r - r0 < 64) ? p0.r - r0 : 64;
    p.A_base = p0.A_base + (int64_t)r0 * p0.np;
    p.U_base = p0.U_base + (int64_t)r0 * p0.d;
    p.Sxb = p0.Sxb + r0;
    p.cross_only = r0 > 0;
    if (r0 > 0) { p.Kt = nullptr; p.A_absmax = nullptr; }   // (mean and row maxima come from the first launch)
    const int rc = posterior_blocks_fwd_chunk(p, st);
    if (rc) return rc;
  }
  return 0;
}

static int posterior_blocks_fwd_chunk(const BlocksParams& p, cudaStream_t st) {
  const int qt_ = (p.q + 7) / 8, rt_ = (p.r + 7) / 8;
  if (qt_ == 1) {
    if (rt_ == 0) return launch_blocks_fwd<1, 0>(p, st);
    if (rt_ == 1) return launch_blocks_fwd<1, 1>(p, st);
    if (rt_ == 2) return launch_blocks_fwd<1, 2>(p, st);
    if (rt_ <= 4) return launch_blocks_fwd<1, 4>(p, st);
    if (rt_ <= 8) return launch_blocks_fwd<1, 8>(p, st);
    return MCACQ_ELIMIT;
  }
  if (qt_ == 2) return posterior_blocks_fwd_q2(p, st);
  if (qt_ <= 4) return posterior_blocks_fwd_q4(p, st);
  return MCACQ_ELIMIT;
}

}  // namespace mcacq
