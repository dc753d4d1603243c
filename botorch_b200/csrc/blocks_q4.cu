// Forward posterior blocks, 17 <= q <= 32, r <= 32 (see blocks.cu).
#include "blocks_fwd_impl.cuh"

namespace mcacq {

int posterior_blocks_fwd_q4r8(const BlocksParams& p, cudaStream_t st);

int posterior_blocks_fwd_q4(const BlocksParams& p, cudaStream_t st) {
  const int rt_ = (p.r + 7) / 8;
  if (rt_ == 0) return launch_blocks_fwd<4, 0>(p, st);
  if (rt_ <= 4) return launch_blocks_fwd<4, 4>(p, st);
  if (rt_ <= 8) return posterior_blocks_fwd_q4r8(p, st);
  return MCACQ_ELIMIT;
}

}  // namespace mcacq
