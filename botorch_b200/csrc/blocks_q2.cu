// Forward posterior blocks, 9 <= q <= 16 (see blocks.cu).
#include "blocks_fwd_impl.cuh"

namespace mcacq {

int posterior_blocks_fwd_q2(const BlocksParams& p, cudaStream_t st) {
  const int rt_ = (p.r + 7) / 8;
  if (rt_ == 0) return launch_blocks_fwd<2, 0>(p, st);
  if (rt_ <= 2) return launch_blocks_fwd<2, 2>(p, st);
  if (rt_ <= 4) return launch_blocks_fwd<2, 4>(p, st);
  if (rt_ <= 8) return launch_blocks_fwd<2, 8>(p, st);
  return MCACQ_ELIMIT;
}

}  // namespace mcacq
