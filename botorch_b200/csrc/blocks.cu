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

int posterior_blocks_fwd(const BlocksParams& p, cudaStream_t st) {
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
