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


// ---- balanced radix-256 digits (int8 contraction mode) ----------------------------------------------------------------
// X (|X| <= 2^(8G-2)) = sum_i d_i 256^i with d_i in [-128, 127].  Adding 128 to every byte position turns the carry
// chain into ONE 64-bit addition: the bytes of Y = X + 0x80..80 are d_i + 128, i.e. d_i = byte_i(Y) ^ 0x80 as int8.
__device__ __forceinline__ unsigned long long balanced_bytes(long long X) {
  return ((unsigned long long)X + 0x0080808080808080ull) ^ 0x0080808080808080ull;
}
// byte `i` (0 = least significant digit) of four packed values -> one 32-bit word (4 consecutive int8 outputs)
__device__ __forceinline__ unsigned pack_digit4(const unsigned long long (&Y)[4], int i) {
  const int sh = 8 * i;
  return (unsigned)((Y[0] >> sh) & 0xFF) | ((unsigned)((Y[1] >> sh) & 0xFF) << 8) | ((unsigned)((Y[2] >> sh) & 0xFF) << 16) |
         ((unsigned)((Y[3] >> sh) & 0xFF) << 24);
}

}  // namespace mcacq
