// Shared pieces of the posterior-block kernels (blocks.cu: forward, blocks_bwd.cu: backward; two translation units so that
// the template instantiations compile in parallel).
#pragma once
#include <cstdlib>

#include "common.cuh"
#include "params.cuh"

namespace mcacq {

__device__ __forceinline__ void dmma884b(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void load4(const double* p, bool ok, double (&v)[4]) {
  if (ok) {
    double2 a = *reinterpret_cast<const double2*>(p);
    double2 b = *reinterpret_cast<const double2*>(p + 2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  } else {
    v[0] = v[1] = v[2] = v[3] = 0.0;
  }
}


template <int N>
__device__ __forceinline__ void loadn(const double* p, bool ok, double (&v)[N]) {
  static_assert(N == 2 || N == 4, "fragment width");
  if (ok) {
    double2 a = *reinterpret_cast<const double2*>(p);
    v[0] = a.x; v[1] = a.y;
    if (N == 4) {
      double2 b = *reinterpret_cast<const double2*>(p + 2);
      v[2] = b.x; v[3] = b.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; i++) v[i] = 0.0;
  }
}

constexpr int BLK_WARPS = 4;

#define MCACQ_DISPATCH_QT_RT(FN, p, st)                                   \
  do {                                                                    \
    const int qt_ = (p.q + 7) / 8, rt_ = (p.r + 7) / 8;                   \
    if (qt_ == 1) {                                                       \
      if (rt_ == 0) return FN<1, 0>(p, st);                               \
      if (rt_ == 1) return FN<1, 1>(p, st);                               \
      if (rt_ == 2) return FN<1, 2>(p, st);                               \
      if (rt_ <= 4) return FN<1, 4>(p, st);                               \
      if (rt_ <= 8) return FN<1, 8>(p, st);                               \
    } else if (qt_ == 2) {                                                \
      if (rt_ == 0) return FN<2, 0>(p, st);                               \
      if (rt_ <= 2) return FN<2, 2>(p, st);                               \
      if (rt_ <= 4) return FN<2, 4>(p, st);                               \
      if (rt_ <= 8) return FN<2, 8>(p, st);                               \
    } else if (qt_ <= 4) {                                                \
      if (rt_ == 0) return FN<4, 0>(p, st);                               \
      if (rt_ <= 4) return FN<4, 4>(p, st);                               \
      if (rt_ <= 8) return FN<4, 8>(p, st);                               \
    }                                                                     \
    return MCACQ_ELIMIT;                                                  \
  } while (0)

}  // namespace mcacq
