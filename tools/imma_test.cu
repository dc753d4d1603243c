// Bring-up test for the tcgen05 int8 path: one 128 x 64 tile, C = A * B^T (A: M x K, B: N x K, both K-major int8).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/imma_test tools/imma_test.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  return (EncodeTiledFn)fn;
}

static CUtensorMap make_map(EncodeTiledFn enc, void* ptr, uint64_t rows, uint64_t kbytes, uint64_t pitch, uint32_t box_rows, uint32_t box_k) {
  CUtensorMap m;
  cuuint64_t dims[2] = {kbytes, rows};
  cuuint64_t strides[1] = {pitch};
  cuuint32_t box[2] = {box_k, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); exit(1); }
  return m;
}

constexpr int BM = 128, BN = 64, BK = 64;  // BK bytes (int8) = one SWIZZLE_64B row
constexpr int STAGES = 2;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT_LOOP;\nDONE:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(const void* smem_ptr) {
  // K-major, SWIZZLE_64B: SBO = 8 rows * 64 B = 512 B; version 1; layout type 4
  uint64_t d = 0;
  d |= (uint64_t)((smem_u32(smem_ptr) & 0x3FFFF) >> 4);
  d |= (uint64_t)0 << 16;                 // LBO (unused for swizzled K-major)
  d |= (uint64_t)(512 >> 4) << 32;        // SBO
  d |= (uint64_t)1 << 46;                 // version
  d |= (uint64_t)4 << 61;                 // SWIZZLE_64B
  return d;
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(128, 1)
imma_tile_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int K, int32_t* C, int ldc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                           // STAGES x (128 x 64)
  uint8_t* sB = smem + STAGES * BM * BK;        // STAGES x (64 x 64)
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], done_bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  const int nk = K / BK;
  // instruction descriptor: D = S32 (2 << 4), A = B = signed int8 (1 << 7, 1 << 10), K-major both, N >> 3 at bit 17, M >> 4 at bit 24
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

  if (warp == 0 && lane == 0) {
    // ---- TMA producer
    for (int it = 0; it < nk; it++) {
      const int s = it % STAGES;
      if (it >= STAGES) mbar_wait(&empty_bar[s], ((it / STAGES) - 1) & 1);
      mbar_expect_tx(&full_bar[s], BM * BK + BN * BK);
      tma_load_2d(sA + s * BM * BK, &mapA, &full_bar[s], it * BK, 0);
      tma_load_2d(sB + s * BN * BK, &mapB, &full_bar[s], it * BK, 0);
    }
  } else if (warp == 1 && lane == 0) {
    // ---- MMA issuer
    for (int it = 0; it < nk; it++) {
      const int s = it % STAGES;
      mbar_wait(&full_bar[s], (it / STAGES) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint64_t adesc = make_desc(sA + s * BM * BK);
      const uint64_t bdesc = make_desc(sB + s * BN * BK);
#pragma unroll
      for (int k = 0; k < BK / 32; k++) {
        // advance 32 bytes along K inside the swizzle atom: +2 in the (>>4) start-address field
        umma_i8(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
      }
      umma_commit(&empty_bar[s]);
    }
    umma_commit(&done_bar);
  }
  __syncthreads();
  // ---- epilogue: all 4 warps, thread i <-> TMEM lane i (row i)
  mbar_wait(&done_bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[64];
  const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
  for (int c = 0; c < 64; c += 16) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[c+0]), "=r"(v[c+1]), "=r"(v[c+2]), "=r"(v[c+3]), "=r"(v[c+4]), "=r"(v[c+5]), "=r"(v[c+6]), "=r"(v[c+7]),
                   "=r"(v[c+8]), "=r"(v[c+9]), "=r"(v[c+10]), "=r"(v[c+11]), "=r"(v[c+12]), "=r"(v[c+13]), "=r"(v[c+14]), "=r"(v[c+15])
                 : "r"(taddr + c));
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int c = 0; c < 64; c++) C[(size_t)tid * ldc + c] = (int32_t)v[c];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64) : "memory");
}

__global__ void ref_kernel(const int8_t* A, const int8_t* B, int K, int32_t* C) {
  int i = blockIdx.x, j = threadIdx.x;
  int s = 0;
  for (int k = 0; k < K; k++) s += (int)A[(size_t)i * K + k] * (int)B[(size_t)j * K + k];
  C[i * 64 + j] = s;
}

int main() {
  const int K = 512;
  std::vector<int8_t> hA(BM * K), hB(BN * K);
  srand(1);
  for (auto& x : hA) x = (int8_t)(rand() % 255 - 127);
  for (auto& x : hB) x = (int8_t)(rand() % 255 - 127);
  int8_t *dA, *dB; int32_t *dC, *dR;
  CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size())); CK(cudaMalloc(&dC, BM * BN * 4)); CK(cudaMalloc(&dR, BM * BN * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dC, 0xff, BM * BN * 4));
  EncodeTiledFn enc = get_encode();
  CUtensorMap mA = make_map(enc, dA, BM, K, K, BM, BK);
  CUtensorMap mB = make_map(enc, dB, BN, K, K, BN, BK);
  size_t smem = STAGES * (BM * BK + BN * BK) + 1024;
  CK(cudaFuncSetAttribute(imma_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  imma_tile_kernel<<<1, 128, smem>>>(mA, mB, K, dC, BN);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  ref_kernel<<<BM, BN>>>(dA, dB, K, dR);
  CK(cudaDeviceSynchronize());
  std::vector<int32_t> hC(BM * BN), hR(BM * BN);
  CK(cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hR.data(), dR, hR.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int i = 0; i < BM * BN; i++) if (hC[i] != hR[i]) { if (bad < 8) printf("mismatch at (%d,%d): got %d want %d\n", i / BN, i % BN, hC[i], hR[i]); bad++; }
  printf("imma tile test: %d mismatches of %d\n", bad, BM * BN);
  return bad != 0;
}
