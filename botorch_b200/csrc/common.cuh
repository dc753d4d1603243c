// Shared device/host helpers for the mcacq_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/mcacq_b200.h"
#include "fast_math.cuh"

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
    return outputscale * fm_exp(-0.5 * sq);
  } else {
    // gpytorch MaternKernel nu=2.5: dist = sqrt(clamp_min(sq, 1e-30))
    double rho = sqrt(fmax(sq, 1e-30));
    double s5r = 2.23606797749978969641 * rho;
    return outputscale * (1.0 + s5r + (5.0 / 3.0) * rho * rho) * fm_exp(-s5r);
  }
}

__device__ __forceinline__ double kernel_dfactor(int kernel_id, double outputscale, double sq) {
  if (kernel_id == MCACQ_KERNEL_RBF) {
    return -outputscale * fm_exp(-0.5 * sq);
  } else {
    // autograd through sqrt(clamp_min(sq, 1e-30)): zero gradient below the clamp
    if (!(sq > 1e-30)) return 0.0;
    double rho = sqrt(sq);
    double s5r = 2.23606797749978969641 * rho;
    return -(5.0 / 3.0) * outputscale * (1.0 + s5r) * fm_exp(-s5r);
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

// All eight digits of four packed values: emit(i, word) receives bytes i of Y[0..3] (4 consecutive int8 outputs of slice
// digit i), most significant half first.  A 4 x 8 byte transpose in 16 PRMT instead of ~12 shift / mask / or operations per
// digit word; each word is handed to `emit` as soon as it exists, so at most four of them are live.
template <typename Emit>
__device__ __forceinline__ void digits4(const unsigned long long (&Y)[4], Emit emit) {
#pragma unroll
  for (int h = 1; h >= 0; h--) {
    const unsigned w0 = (unsigned)(Y[0] >> (32 * h)), w1 = (unsigned)(Y[1] >> (32 * h));
    const unsigned w2 = (unsigned)(Y[2] >> (32 * h)), w3 = (unsigned)(Y[3] >> (32 * h));
    const unsigned t0 = __byte_perm(w0, w1, 0x5140), t1 = __byte_perm(w2, w3, 0x5140);
    emit(4 * h + 0, __byte_perm(t0, t1, 0x5410));
    emit(4 * h + 1, __byte_perm(t0, t1, 0x7632));
    const unsigned t2 = __byte_perm(w0, w1, 0x7362), t3 = __byte_perm(w2, w3, 0x7362);
    emit(4 * h + 2, __byte_perm(t2, t3, 0x5410));
    emit(4 * h + 3, __byte_perm(t2, t3, 0x7632));
  }
}

// Two groups at once (8 consecutive outputs): emit(i, word of Ya, word of Yb).
template <typename Emit>
__device__ __forceinline__ void digits4x2(const unsigned long long (&Ya)[4], const unsigned long long (&Yb)[4], Emit emit) {
#pragma unroll
  for (int h = 1; h >= 0; h--) {
    const unsigned a0 = (unsigned)(Ya[0] >> (32 * h)), a1 = (unsigned)(Ya[1] >> (32 * h));
    const unsigned a2 = (unsigned)(Ya[2] >> (32 * h)), a3 = (unsigned)(Ya[3] >> (32 * h));
    const unsigned b0 = (unsigned)(Yb[0] >> (32 * h)), b1 = (unsigned)(Yb[1] >> (32 * h));
    const unsigned b2 = (unsigned)(Yb[2] >> (32 * h)), b3 = (unsigned)(Yb[3] >> (32 * h));
    {
      const unsigned s0 = __byte_perm(a0, a1, 0x5140), s1 = __byte_perm(a2, a3, 0x5140);
      const unsigned u0 = __byte_perm(b0, b1, 0x5140), u1 = __byte_perm(b2, b3, 0x5140);
      emit(4 * h + 0, __byte_perm(s0, s1, 0x5410), __byte_perm(u0, u1, 0x5410));
      emit(4 * h + 1, __byte_perm(s0, s1, 0x7632), __byte_perm(u0, u1, 0x7632));
    }
    {
      const unsigned s2 = __byte_perm(a0, a1, 0x7362), s3 = __byte_perm(a2, a3, 0x7362);
      const unsigned u2 = __byte_perm(b0, b1, 0x7362), u3 = __byte_perm(b2, b3, 0x7362);
      emit(4 * h + 2, __byte_perm(s2, s3, 0x5410), __byte_perm(u2, u3, 0x5410));
      emit(4 * h + 3, __byte_perm(s2, s3, 0x7632), __byte_perm(u2, u3, 0x7632));
    }
  }
}

}  // namespace mcacq
