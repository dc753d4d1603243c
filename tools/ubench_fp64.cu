// Micro-benchmark: raw FP64 pipe rates on B200 (DMMA.8x8x4 vs DFMA), register-only loops.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fp64 ubench_fp64.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

template <int NACC>
__global__ void k_dmma(double* out, const double* in, int iters) {
  double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
  double c[NACC][2];
#pragma unroll
  for (int j = 0; j < NACC; j++) { c[j][0] = 0; c[j][1] = 0; }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < NACC; j++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < NACC; j++) s += c[j][0] + c[j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dfma(double* out, const double* in, int iters) {
  double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
  double c[NACC];
#pragma unroll
  for (int j = 0; j < NACC; j++) c[j] = j;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < NACC; j++) c[j] = fma(a, c[j], b);
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < NACC; j++) s += c[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
  double *in, *out; cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096); cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  int iters = 20000;
  for (int warps : {1, 2, 4, 8, 16, 32}) {
    for (int ctas_per_sm : {1, 2}) {
      int grid = sms * ctas_per_sm, block = warps * 32;
      if (block * ctas_per_sm > 2048) continue;
      float ms = timeit([&] { k_dmma<8><<<grid, block>>>(out, in, iters); });
      double flops = 2.0 * 256 * 8 * (double)iters * warps * grid;
      float ms2 = timeit([&] { k_dfma<8><<<grid, block>>>(out, in, iters); });
      double flops2 = 2.0 * 32 * 8 * (double)iters * warps * grid;
      printf("warps/cta %2d ctas/sm %d : DMMA %.2f TF/s (%.3f ms)   DFMA %.2f TF/s (%.3f ms)\n", warps, ctas_per_sm,
             flops / ms * 1e-9, ms, flops2 / ms2 * 1e-9, ms2);
    }
  }
  // latency: single warp, single accumulator chain
  {
    float ms = timeit([&] { k_dmma<1><<<1, 32>>>(out, in, iters * 10); });
    printf("DMMA dependent chain: %.1f ns per mma\n", ms * 1e6 / (iters * 10));
    float ms2 = timeit([&] { k_dfma<1><<<1, 32>>>(out, in, iters * 10); });
    printf("DFMA dependent chain: %.1f ns per fma\n", ms2 * 1e6 / (iters * 10));
    for (int w : {1, 2, 4}) {
      float ms3 = timeit([&] { k_dmma<8><<<1, 32 * w>>>(out, in, iters); });
      printf("DMMA 8 indep acc, %d warp(s) one SM: %.1f ns per mma => %.1f FMA/ns/SM\n", w, ms3 * 1e6 / (iters * 8.0), 256.0 * w * iters * 8 / (ms3 * 1e6));
    }
  }
  return 0;
}
