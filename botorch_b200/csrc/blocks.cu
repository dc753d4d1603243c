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

// Baselines beyond 64 points: the accumulators of the cross-Gram live in registers (8 row tiles at most), so the baseline is
// swept in chunks of 64 rows, one launch per chunk.  The first launch also produces mean / Sxx / the row maxima; the others
// only write their columns of Sxb.  Every Sxb entry is the same DMMA sequence over k as in a single-chunk launch.
int posterior_blocks_fwd(const BlocksParams& p0, cudaStream_t st) {
  BlocksParams p = p0;
  p.r_pitch = p0.r;
  p.cross_only = 0;
  if (p0.r <= 64) return posterior_blocks_fwd_chunk(p, st);
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
