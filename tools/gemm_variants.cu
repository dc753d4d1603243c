// Developer tool: benchmark tile-shape variants of the triangular DMMA GEMM (run under gpurun).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I botorch_b200/csrc -o tools/gemm_variants tools/gemm_variants.cu
#include "dgemm_tri.cuh"
#include <cstdio>
#include <vector>
#include <cmath>
namespace mcacq { thread_local int g_launch_count = 0; __global__ void zero_counter_kernel(int* c) { *c = 0; } }
using namespace mcacq;

__global__ void fill(double* p, size_t n, unsigned seed) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) { unsigned x = (unsigned)(i * 2654435761u) ^ seed; x ^= x >> 13; x *= 0x5bd1e995; x ^= x >> 15; p[i] = (x & 0xffff) / 65536.0 - 0.5; }
}
__global__ void triu(double* p, int n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < (size_t)n * n) { int r = i / n, c = i % n; if (r > c) p[i] = 0.0; }
}
__global__ void ref_rows(const double* A, const double* B, double* out, int np, const int64_t* rows, int nrows) {
  int c = blockIdx.x * blockDim.x + threadIdx.x; int r = blockIdx.y;
  if (c < np) { double s = 0; for (int k = 0; k < np; k++) s += A[rows[r] * np + k] * B[(size_t)k * np + c]; out[(size_t)r * np + c] = s; }
}

template <typename F> float timeit(F f, int reps = 3) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize(); float best = 1e9;
  for (int i = 0; i < reps; i++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = fminf(best, ms); }
  return best;
}

int main(int argc, char** argv) {
  int64_t M = argc > 1 ? atoll(argv[1]) : 65536; int np = argc > 2 ? atoi(argv[2]) : 4096;
  double *A, *B, *C, *Cref; int* counter; int64_t* rows;
  cudaMalloc(&A, M * np * 8); cudaMalloc(&B, (size_t)np * np * 8); cudaMalloc(&C, M * np * 8); cudaMalloc(&counter, 256);
  const int NR = 16; cudaMalloc(&Cref, (size_t)NR * np * 8); cudaMalloc(&rows, NR * 8);
  fill<<<(M * np + 255) / 256, 256>>>(A, M * np, 1); fill<<<((size_t)np * np + 255) / 256, 256>>>(B, (size_t)np * np, 2);
  triu<<<((size_t)np * np + 255) / 256, 256>>>(B, np);
  std::vector<int64_t> hr(NR); for (int i = 0; i < NR; i++) hr[i] = (i * 7919 + 13) % M; hr[NR - 1] = M - 1;
  cudaMemcpy(rows, hr.data(), NR * 8, cudaMemcpyHostToDevice);
  ref_rows<<<dim3((np + 127) / 128, NR), 128>>>(A, B, Cref, np, rows, NR);
  std::vector<double> href((size_t)NR * np), hc(np);
  cudaMemcpy(href.data(), Cref, (size_t)NR * np * 8, cudaMemcpyDeviceToHost);
  auto check = [&](const char* name, float ms) {
    double maxerr = 0;
    for (int i = 0; i < NR; i++) { cudaMemcpy(hc.data(), C + hr[i] * np, np * 8, cudaMemcpyDeviceToHost);
      for (int c = 0; c < np; c++) maxerr = fmax(maxerr, fabs(hc[c] - href[(size_t)i * np + c])); }
    double fl = (double)M * np * (np + 1);
    printf("%-28s %8.3f ms  %6.2f TF/s alg  maxerr %.2e\n", name, ms, fl / ms * 1e-9, maxerr); fflush(stdout);
  };
#define RUN(NAME, ...) { cudaMemset(C, 0, M * np * 8); float ms = timeit([&] { launch_dgemm_tri<__VA_ARGS__>(MCACQ_TRI_UPPER, M, np, A, B, C, counter, 0); }); \
    cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("%s: CUDA error %s\n", NAME, cudaGetErrorString(e)); return 1; } check(NAME, ms); }
  RUN("64x64 w2x2 s4 occ3 bk16", 64, 64, 2, 2, 4, 3, 16)
  RUN("64x64 w2x2 s3 occ3 bk16", 64, 64, 2, 2, 3, 3, 16)
  RUN("64x64 w2x2 s3 occ4 bk16", 64, 64, 2, 2, 3, 4, 16)
  RUN("64x64 w2x2 s2 occ3 bk32", 64, 64, 2, 2, 2, 3, 32)
  RUN("64x64 w2x2 s3 occ2 bk32", 64, 64, 2, 2, 3, 2, 32)
  RUN("128x64 w4x2 s3 occ2 bk16", 128, 64, 4, 2, 3, 2, 16)
  RUN("128x64 w4x2 s4 occ2 bk16", 128, 64, 4, 2, 4, 2, 16)
  RUN("128x64 w4x2 s2 occ2 bk32", 128, 64, 4, 2, 2, 2, 32)
  RUN("64x128 w2x4 s2 occ2 bk32", 64, 128, 2, 4, 2, 2, 32)
  RUN("128x64 w2x2 s3 occ2 bk16", 128, 64, 2, 2, 3, 2, 16)
  RUN("64x64 w1x2 s4 occ3 bk16", 64, 64, 1, 2, 4, 3, 16)
  RUN("64x64 w1x2 s4 occ4 bk16", 64, 64, 1, 2, 4, 4, 16)
  RUN("32x64 w1x2 s4 occ6 bk16", 32, 64, 1, 2, 4, 6, 16)
  RUN("64x32 w2x1 s4 occ6 bk16", 64, 32, 2, 1, 4, 6, 16)
  return 0;
}
