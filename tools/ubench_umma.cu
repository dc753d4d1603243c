// Microbenchmark (developer tool, run under gpurun): issue rate of tcgen05.mma with BOTH operands in shared memory
// ("SS"), M = 128 per CTA, K = 32 bytes per instruction, as a function of kind (i8 / f8f6f4 / f16), N, swizzle width and
// cta_group.  One CTA (or CTA pair) per SM, operands are whatever the shared memory holds; one thread issues `iters`
// accumulating instructions back to back and waits for their completion through tcgen05.commit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ubench_umma tools/ubench_umma.cu
// Answers: what does the tensor pipe sustain for the instruction shapes ozaki_imma_kernel issues, independent of TMA,
// barriers and the epilogue?
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra W_DONE;\nbra W_LOOP;\nW_DONE:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(const void* smem_ptr, int bk) {
  // K-major operand tile, SWIZZLE_64B (bk = 64) or SWIZZLE_128B (bk = 128): SBO = 8 rows * bk bytes; descriptor version 1
  uint64_t d = (uint64_t)((smem_u32(smem_ptr) & 0x3FFFF) >> 4);
  d |= (uint64_t)((8 * bk) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(bk == 128 ? 2 : 4) << 61;
  return d;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// KIND: 0 = kind::i8 (S8 x S8 -> S32), 1 = kind::f8f6f4 (e4m3 -> fp32), 2 = kind::f16 (bf16 -> fp32)
template <int KIND, int CG, int uniform>
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(int N, int iters, int bk, int same_acc, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  uint32_t crank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  // deterministic, non-trivial operand bytes
  for (int i = tid; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u * (uint32_t)(i % 7);
  if (tid == 0) { mbar_init(&done_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  // A: 128 rows x bk bytes at offset 0 (k-steps of 32 bytes inside);  B: N (or N/2) rows x bk bytes behind it
  const uint8_t* sA = smem;
  const uint8_t* sB = smem + 128 * bk;
  const int ksteps = bk / 32;
  uint32_t idesc;
  if (KIND == 0) idesc = (2u << 4) | (1u << 7) | (1u << 10);
  else if (KIND == 1) idesc = (1u << 4);                       // D = fp32, A = B = e4m3 (format 0)
  else idesc = (1u << 4) | (1u << 7) | (1u << 10);             // D = fp32, A = B = bf16
  idesc |= ((uint32_t)(N >> 3) << 17) | ((uint32_t)((CG == 2 ? 256 : 128) >> 4) << 24);
  // uniform = 1: the whole warp runs the issue loop in warp-uniform control flow and only the instruction itself sits
  // under elect.sync (the operands stay in uniform registers); uniform = 0: one thread in a divergent branch (ptxas then
  // wraps every UTCIMMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY uniformisation loop)
  const bool issuer = uniform ? (warp == 1 && crank == 0) : (tid == 32 && crank == 0);
  if (issuer) {
    const uint64_t ad0 = make_desc(sA, bk), bd0 = make_desc(sB, bk);
    const bool lead = uniform ? elect_one() : true;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
      const int k = it & (ksteps - 1);
      if (lead) {
      // same_acc = 1: every instruction accumulates into the same N columns (a GEMM main loop);
      // same_acc = 0: alternate between two column ranges (independent accumulators, N <= 256)
      const uint32_t dcol = same_acc ? 0u : (uint32_t)((it & 1) * 256);
      const uint64_t ad = ad0 + 2 * k, bd = bd0 + 2 * k;
      if (KIND == 0) {
        if (CG == 1) asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_base + dcol), "l"(ad), "l"(bd), "r"(idesc), "r"(it > 1 ? 1u : 0u) : "memory");
        else asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_base + dcol), "l"(ad), "l"(bd), "r"(idesc), "r"(it > 1 ? 1u : 0u) : "memory");
      } else if (KIND == 1) {
        if (CG == 1) asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_base + dcol), "l"(ad), "l"(bd), "r"(idesc), "r"(it > 1 ? 1u : 0u) : "memory");
        else asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_base + dcol), "l"(ad), "l"(bd), "r"(idesc), "r"(it > 1 ? 1u : 0u) : "memory");
      } else {
        if (CG == 1) asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_base + dcol), "l"(ad), "l"(bd), "r"(idesc), "r"(it > 1 ? 1u : 0u) : "memory");
        else asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_base + dcol), "l"(ad), "l"(bd), "r"(idesc), "r"(it > 1 ? 1u : 0u) : "memory");
      }
      }
    }
    if (lead) {
    if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done_bar)) : "memory");
    else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done_bar)) : "memory");
    }
    mbar_wait(&done_bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && lead) cycles[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync();
  if (warp == 0) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

template <int KIND, int CG, int uniform = 1>
static void run(const char* name, int N, int bk, int same_acc, int sms, long long* d_cycles) {
  const int iters = 4096;
  auto kern = umma_rate_kernel<KIND, CG, uniform>;
  const size_t smem = 201 * 1024 + 1024;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  cfg.blockDim = dim3(128);
  cfg.gridDim = dim3(CG == 2 ? sms / 2 * 2 : sms);
  cfg.dynamicSmemBytes = smem;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CG == 2 ? 1 : 0;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; rep++) {
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, kern, N, iters, bk, same_acc, d_cycles));
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
  }
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  long long cyc = 0;
  CK(cudaMemcpy(&cyc, d_cycles, 8, cudaMemcpyDeviceToHost));
  const double per = (double)cyc / iters;
  const double kelem = (KIND == 2) ? 16.0 : 32.0;
  const double macs_per_clk_sm = 128.0 * N * kelem / per;   // per SM (each CTA of a pair owns 128 rows)
  printf("%-10s %s cta_group::%d N=%3d swizzle=%3dB same_acc=%d : %7.1f clk/instr  %7.0f MAC/clk/SM  (%.2f clk/column)  kernel %.3f ms\n",
         name, uniform ? "uniform-issue" : "1-thread-issue", CG, N, bk, same_acc, per, macs_per_clk_sm, per / N, ms);
}

int main() {
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  long long* d_cycles;
  CK(cudaMalloc(&d_cycles, 8));
  run<0, 1, 0>("i8", 256, 64, 1, sms, d_cycles);
  run<0, 1, 0>("i8", 80, 64, 1, sms, d_cycles);
  for (int bk : {64, 128}) {
    for (int N : {64, 80, 128, 160, 240, 256}) {
      run<0, 1>("i8", N, bk, 1, sms, d_cycles);
    }
    run<0, 1>("i8", 256, bk, 0, sms, d_cycles);
    run<1, 1>("f8f6f4", 256, bk, 1, sms, d_cycles);
    run<2, 1>("f16(bf16)", 256, bk, 1, sms, d_cycles);
    for (int N : {64, 128, 160, 256}) run<0, 2>("i8", N, bk, 1, sms, d_cycles);
    run<1, 2>("f8f6f4", 256, bk, 1, sms, d_cycles);
    run<2, 2>("f16(bf16)", 256, bk, 1, sms, d_cycles);
  }
  return 0;
}
