// Shared device/host helpers for the mcacq_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/mcacq_b200.h"

#define MCACQ_CUDA_CHECK_LAUNCH()                      \
  do {                                                 \
    cudaError_t _e = cudaGetLastError();               \
    if (_e != cudaSuccess) return (int)_e;             \
  } while (0)

namespace mcacq {

extern thread_local int g_launch_count;
inline void count_launch(int k = 1) { g_launch_count += k; }

__host__ __device__ inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// ---- covariance function of the scaled distance --------------------------------------------
// value: k(rho^2); deriv: g such that dk/du_i = g * (u_i - u_j)  (g = (dk/drho) / rho).
__device__ __forceinline__ double kernel_value(int kernel_id, double outputscale, double sq) {
  if (kernel_id == MCACQ_KERNEL_RBF) {
    return outputscale * exp(-0.5 * sq);
  } else {
    // gpytorch MaternKernel nu=2.5: dist = sqrt(clamp_min(sq, 1e-30))
    double rho = sqrt(fmax(sq, 1e-30));
    double s5r = 2.23606797749978969641 * rho;
    return outputscale * (1.0 + s5r + (5.0 / 3.0) * rho * rho) * exp(-s5r);
  }
}

__device__ __forceinline__ double kernel_dfactor(int kernel_id, double outputscale, double sq) {
  if (kernel_id == MCACQ_KERNEL_RBF) {
    return -outputscale * exp(-0.5 * sq);
  } else {
    // autograd through sqrt(clamp_min(sq, 1e-30)): zero gradient below the clamp
    if (!(sq > 1e-30)) return 0.0;
    double rho = sqrt(sq);
    double s5r = 2.23606797749978969641 * rho;
    return -(5.0 / 3.0) * outputscale * (1.0 + s5r) * exp(-s5r);
  }
}

}  // namespace mcacq
