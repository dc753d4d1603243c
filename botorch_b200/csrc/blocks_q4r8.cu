// Forward posterior blocks, 17 <= q <= 32, 33 <= r <= 64 (see blocks.cu).
#include "blocks_fwd_impl.cuh"

namespace mcacq {

int posterior_blocks_fwd_q4r8(const BlocksParams& p, cudaStream_t st) { return launch_blocks_fwd<4, 8>(p, st); }

}  // namespace mcacq
