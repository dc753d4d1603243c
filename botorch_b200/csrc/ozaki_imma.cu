// FP64-accurate contraction on the INT8 tensor cores (tcgen05 `kind::i8`, TMEM int32 accumulators, TMA operands).
//
//   C[M x N] (fp64) = A[M x K] * B^T,   B given as rows [N x K] (K contiguous),   triangular-aware in K
//
// Ozaki-style error-free splitting: both operands are pre-split into signed 8-bit slices (balanced radix-256 digits)
// with a power-of-two row scale,  A[i][k] = ea[i] * sum_p A_p[i][k] 256^-(p+1),  B[j][k] = fb[j] * sum_q B_q[j][k] 256^-(q+1)
// (`slice_rows_kernel` below).  Every slice product A_p B_q^T is EXACT in int32 (G * K * 128^2 < 2^31), so
//   C[i][j] = ea[i] * fb[j] * sum_{g < G} 256^-(g+2) * S_g[i][j],     S_g = sum_{p+q=g} A_p B_q^T,
// with truncation error ~2^-(8G-2) relative to (row max) x (column max): G = 6 diagonals (46-bit operands) keep the
// posterior variance within ~1e-10, G = 5 ~1e-8 (tools/ozaki_test.py; the first version used 7-bit truncated digits
// and needed G = 7 / 6 for the same accuracy).  This is the same contraction as dgemm_tri.cu
// (gpytorch's `test_train_covar @ covar_cache`), executed at INT8 tensor-core rate instead of the FP64 DMMA pipe.
//
// Kernel structure (one persistent CTA per SM, 6 warps):
//   warp 0 / lane 0 : TMA producer.  Per 64-byte k-block it loads G A-tiles (128 rows x 64 B) and G B-tiles
//                     (BN rows x 64 B), SWIZZLE_64B, into a ring of G * (128 + BN) * 64 B stages (2 at G = 6 / BN = 80,
//                     3 at G = 5 / BN = 96; 128-byte k-blocks, 7-10 % faster per byte, when two such stages fit: G <= 3).
//   warp 1 / lane 0 : MMA issuer.  All G diagonals S_g live in TMEM at once (G x BN columns of int32; BN is the widest
//                     multiple of 16 with G * BN <= 512), so one k-block of operands feeds G(G+1)/2 slice products.
//                     Because S_p .. S_{G-1} are adjacent TMEM column ranges, A_p x [B_0; ..; B_{G-1-p}] is issued as
//                     wide `tcgen05.mma.cta_group::1.kind::i8` instructions (M=128, N<=256, K=32): 9 per K step at
//                     G = 6 instead of 21.
//   warps 2-5       : epilogue.  Thread <-> TMEM lane <-> output row; 16 columns at a time: `tcgen05.ld` all G
//                     diagonals, merge triples exactly in int64, convert with the 2^52+2^51 constant, combine in fp64,
//                     apply the row/column scales, store one 128-byte line per thread.
// Tiles are visited in an L2-friendly static order (groups of 8 row-tiles sweep the column tiles from the longest
// k-range to the shortest); all three roles derive the same sequence from blockIdx, so no tile broadcast is needed.
// Measured limits (profiles/r01_ozaki_int8.md): ~19-20 B/clk of operand bytes per SM and ~1.3 clk per accumulator column
// per UMMA -- equal walls at G = 6.  Experimental variants kept behind environment switches: TMA-multicast clusters
// (MCACQ_OZ_CLUSTER) and a cta_group::2 pair kernel (MCACQ_OZ_CTA2), both numerically identical and slower.
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"

namespace mcacq {

constexpr int OZ_BM = 128;               // BK (bytes = int8 elements) is a template parameter: 64 (SWIZZLE_64B) or 128 (SWIZZLE_128B)
constexpr int OZ_MAX_BN = 256;           // the column-tile width BN is a template parameter: the widest multiple of 16 with G * BN <= 512
                                         // TMEM columns (G = 6: 80, G = 5: 96, G = 4: 128), because the operand bytes an SM has to
                                         // ingest per MAC fall as BM * BN / (BM + BN)
constexpr int OZ_MAX_STAGES = 4;                    // ring depth is chosen at run time: as many 12*G KB stages as fit
constexpr int OZ_MAXG = 7;
constexpr int OZ_GROUP_M = 8;
constexpr int OZ_THREADS = 192;

__device__ __forceinline__ uint32_t oz_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void oz_mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(oz_smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void oz_mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(oz_smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void oz_mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(oz_smem_u32(b)) : "memory");
}
__device__ __forceinline__ void oz_mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nOZ_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra OZ_DONE;\nbra OZ_WAIT;\nOZ_DONE:\n}"
               ::"r"(oz_smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void oz_tma_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(oz_smem_u32(dst)), "l"(map), "r"(oz_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void oz_tma_3d_hint(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                               uint64_t policy) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
               ::"r"(oz_smem_u32(dst)), "l"(map), "r"(oz_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(policy) : "memory");
}
__device__ __forceinline__ void oz_tma_3d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;"
               ::"r"(oz_smem_u32(dst)), "l"(map), "r"(oz_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(mask) : "memory");
}
__device__ __forceinline__ void oz_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(oz_smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ bool oz_elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void oz_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t oz_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

template <int OZ_BK>
__device__ __forceinline__ uint64_t oz_desc(const void* smem_ptr) {
  // K-major operand tile, SWIZZLE_64B / 128B: stride between 8-row groups = 8 * BK bytes; descriptor version 1 (sm_100)
  uint64_t d = (uint64_t)((oz_smem_u32(smem_ptr) & 0x3FFFF) >> 4);
  d |= (uint64_t)((8 * OZ_BK) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(OZ_BK == 128 ? 2 : 4) << 61;
  return d;
}
__device__ __forceinline__ void oz_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void oz_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(oz_smem_u32(bar)) : "memory");
}

struct OzTile { int64_t mt; int nt; int kb0, kb1; };

// t-th CLUSTER tile in the static order (a cluster tile = CM x CN adjacent CTA tiles that share operand loads);
// returns false past the end.  The k-range is the union over the cluster's column tiles (the extra k-blocks of the
// narrower columns multiply stored zeros of the triangular factor).
template <int OZ_BK, int OZ_BN, int CM, int CN>
__device__ __forceinline__ bool oz_tile(int64_t t, int64_t m_tiles, int n_tiles, int k_blocks, int tri_mode, int group_m,
                                        OzTile& o) {
  const int64_t cm_tiles = (m_tiles + CM - 1) / CM;
  const int cn_tiles = (n_tiles + CN - 1) / CN;
  if (t >= cm_tiles * cn_tiles) return false;
  const int GM = (group_m / CM) > 0 ? (group_m / CM) : 1;
  const int64_t group_sz = (int64_t)GM * cn_tiles;
  const int64_t grp = t / group_sz;
  const int64_t m0 = grp * GM;
  const int64_t rows = (cm_tiles - m0 < GM) ? (cm_tiles - m0) : GM;
  const int64_t within = t - grp * group_sz;
  const int nrank = (int)(within / rows);
  o.mt = m0 + within % rows;                                                   // cluster row-tile index
  o.nt = (tri_mode == MCACQ_TRI_LOWER) ? nrank : cn_tiles - 1 - nrank;           // cluster column-tile index
  o.kb0 = 0; o.kb1 = k_blocks;
  if (tri_mode == MCACQ_TRI_UPPER) { int e = ((o.nt + 1) * CN * OZ_BN + OZ_BK - 1) / OZ_BK; o.kb1 = e < k_blocks ? e : k_blocks; }
  else if (tri_mode == MCACQ_TRI_LOWER) { int b = (o.nt * CN * OZ_BN) / OZ_BK; o.kb0 = b < k_blocks ? b : k_blocks; }
  return true;
}

// Epilogue role (warps 2..5) shared by the contraction kernels: thread <-> TMEM lane <-> output row.
template <int OZ_BK, int OZ_BN, int CM, int CN>
__device__ __forceinline__ void oz_epilogue(const CUtensorMap& mapC, int tri_mode, int64_t M, int N, int G,
                                            const double* __restrict__ row_scale, const double* __restrict__ col_scale,
                                            double* __restrict__ C, int64_t ldc, int l2_hints, int group_m, int tma_store,
                                            int64_t m_tiles, int n_tiles, int k_blocks, int64_t cluster_id, int64_t num_clusters,
                                            int rm, int rn, int warp, int lane, int tid, uint32_t tmem_base, uint64_t& acc_full,
                                            uint64_t& acc_empty_ref, double* s_col, uint8_t* epi_stage) {
  uint64_t* acc_full_p = &acc_full;
  uint64_t* acc_empty_p = &acc_empty_ref;
  const int lg = warp & 3;                  // TMEM lane group this warp may access
  const int r_in_tile = lg * 32 + lane;
  const int etid = tid - 64;                // 0..127
  int64_t tile_i = 0;
  OzTile tl;
  // output staging: 2 x 4 KB (32 rows x 128 B, SWIZZLE_128B) per epilogue warp, behind the operand ring
  epi_stage = (uint8_t*)(((uintptr_t)epi_stage + 1023) & ~(uintptr_t)1023);
  uint32_t epi_chunk = 0;
  for (int64_t t = cluster_id; oz_tile<OZ_BK, OZ_BN, CM, CN>(t, m_tiles, n_tiles, k_blocks, tri_mode, group_m, tl); t += num_clusters, tile_i++) {
    const int64_t row = (tl.mt * CM + rm) * OZ_BM + r_in_tile;
    const int col0 = (tl.nt * CN + rn) * OZ_BN;
    // stage the column scales of this tile (epilogue-only named barrier, 128 threads)
    for (int c = etid; c < OZ_BN; c += 128) s_col[c] = (col0 + c < N) ? col_scale[col0 + c] : 0.0;
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const bool have_acc = tl.kb1 > tl.kb0;
    if (have_acc) {
      oz_mbar_wait(acc_full_p, (uint32_t)(tile_i & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const bool row_ok = row < M;
    const double rs = row_ok ? row_scale[row] : 0.0;
    double* dst = C + (row_ok ? row : 0) * ldc + col0;
    const bool vec_ok = (ldc & 1) == 0;
    // 16 output columns at a time: all G diagonals of the chunk are fetched from TMEM, combined in fp64 with the
    // weights 256^-(g+2) and scaled.  Output: the warp's 32 rows x 16 columns are staged in shared memory in the
    // SWIZZLE_128B layout (conflict-free 16-byte stores) and written by ONE TMA store per chunk, i.e. as full 128-byte
    // lines -- a thread-per-row store pattern hits 32 different lines per instruction with half-filled sectors and made
    // the epilogue (which cannot overlap the main loop: all 512 TMEM columns hold accumulators) 23 % of the kernel.
#pragma unroll 1
    for (int c0 = 0; c0 < OZ_BN; c0 += 16) {
      double acc[16];
#pragma unroll
      for (int j = 0; j < 16; j++) acc[j] = 0.0;
      if (have_acc) {
        uint32_t v[OZ_MAXG][16];
#pragma unroll
        for (int g = 0; g < OZ_MAXG; g++) {
          if (g < G) {
            const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(g * OZ_BN + c0);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(v[g][0]), "=r"(v[g][1]), "=r"(v[g][2]), "=r"(v[g][3]), "=r"(v[g][4]), "=r"(v[g][5]), "=r"(v[g][6]),
                           "=r"(v[g][7]), "=r"(v[g][8]), "=r"(v[g][9]), "=r"(v[g][10]), "=r"(v[g][11]), "=r"(v[g][12]),
                           "=r"(v[g][13]), "=r"(v[g][14]), "=r"(v[g][15])
                         : "r"(taddr));
          }
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        // Three neighbouring diagonals are merged exactly in int64 (|S_g| < 2^31, so the sum stays below 2^48) and
        // converted with the 2^52 + 2^51 magic constant: one DADD + one DFMA per triple instead of an I2F.F64 (16/clk/SM)
        // and a DFMA per diagonal.
        double w = 1.0 / 4294967296.0;  // 256^-4: weight of the last member of the triple (S_0, S_1, S_2)
#pragma unroll
        for (int g0 = 0; g0 < OZ_MAXG; g0 += 3) {
          if (g0 < G) {
#pragma unroll
            for (int j = 0; j < 16; j++) {
              const long long s0 = (long long)(int32_t)v[g0][j];
              const long long s1 = (g0 + 1 < OZ_MAXG && g0 + 1 < G) ? (long long)(int32_t)v[g0 + 1 < OZ_MAXG ? g0 + 1 : g0][j] : 0ll;
              const long long s2 = (g0 + 2 < OZ_MAXG && g0 + 2 < G) ? (long long)(int32_t)v[g0 + 2 < OZ_MAXG ? g0 + 2 : g0][j] : 0ll;
              const long long t = s0 * 65536ll + s1 * 256ll + s2;
              const double d = __longlong_as_double(0x4338000000000000ll + t) - 6755399441055744.0;
              acc[j] = fma(w, d, acc[j]);
            }
            w *= (1.0 / 16777216.0);
          }
        }
      }
      if (tma_store) {
        if (col0 + c0 < N) {   // warp-uniform
          uint8_t* sbuf = epi_stage + (size_t)((lg * 2 + (epi_chunk & 1)) * 4096);
          // the store issued two chunks ago read from this buffer: it must have finished reading before we overwrite
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const double2 o = make_double2(acc[j] * rs * s_col[c0 + j], acc[j + 1] * rs * s_col[c0 + j + 1]);
            *reinterpret_cast<double2*>(sbuf + lane * 128 + ((((j >> 1) ^ (lane & 7))) << 4)) = o;
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                         ::"l"(&mapC), "r"(oz_smem_u32(sbuf)), "r"(col0 + c0), "r"((int)((tl.mt * CM + rm) * OZ_BM + lg * 32)) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          epi_chunk++;
        }
      } else if (row_ok && col0 + c0 < N) {
        if (col0 + c0 + 16 <= N && vec_ok) {
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const double2 o = make_double2(acc[j] * rs * s_col[c0 + j], acc[j + 1] * rs * s_col[c0 + j + 1]);
            if (l2_hints) __stcs(reinterpret_cast<double2*>(dst + c0 + j), o);  // streamed: consumed by a later kernel
            else *reinterpret_cast<double2*>(dst + c0 + j) = o;
          }
        } else {
          for (int j = 0; j < 16; j++) if (col0 + c0 + j < N) dst[c0 + j] = acc[j] * rs * s_col[c0 + j];
        }
      }
    }
    // hand the accumulators back (also for an empty k-range, which cannot happen for the triangular factors, so
    // that the MMA thread's phase bookkeeping stays aligned)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    if (lane == 0) oz_mbar_arrive(acc_empty_p);
    asm volatile("bar.sync 1, 128;" ::: "memory");  // s_col reuse
  }
  if (tma_store && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores complete before exit
}

template <int OZ_BK, int OZ_BN, int CM, int CN>
__global__ void __launch_bounds__(OZ_THREADS, 1)
ozaki_imma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                  const __grid_constant__ CUtensorMap mapC, int tri_mode,
                  int64_t M, int N, int K, int G, int stages, const double* __restrict__ row_scale,
                  const double* __restrict__ col_scale, double* __restrict__ C, int64_t ldc, int l2_hints, int group_m, int dbg,
                  int tma_store) {
  // dbg (developer switches, MCACQ_OZ_DEBUG): bit 0 = do not issue the UMMAs (pure TMA ingest rate), bit 1 = do not issue
  // the TMA loads (pure MMA + epilogue rate on stale shared memory), bit 2 = one TMA box per operand spanning all G slices.
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int OZ_A_TILE = OZ_BM * OZ_BK, OZ_B_TILE = OZ_BN * OZ_BK;
  const int stage_bytes = G * (OZ_A_TILE + OZ_B_TILE);
  __shared__ __align__(8) uint64_t full_bar[OZ_MAX_STAGES], empty_bar[OZ_MAX_STAGES], acc_full, acc_empty;
  __shared__ uint32_t tmem_base_s;
  __shared__ double s_col[OZ_MAX_BN];
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform by construction (and visibly so to ptxas)
  const int64_t m_tiles = (M + OZ_BM - 1) / OZ_BM;
  const int n_tiles = (N + OZ_BN - 1) / OZ_BN;
  const int k_blocks = (K + OZ_BK - 1) / OZ_BK;

  // position inside the cluster: CTAs of one cluster row share the A tiles, CTAs of one cluster column the B tiles
  constexpr int CSIZE = CM * CN;
  const uint32_t crank = (CSIZE > 1) ? oz_cluster_rank() : 0u;
  const int rm = (int)crank / CN, rn = (int)crank % CN;
  const uint16_t row_mask = (uint16_t)(((1u << CN) - 1u) << (rm * CN));
  uint16_t col_mask = 0;
#pragma unroll
  for (int i = 0; i < CM; i++) col_mask |= (uint16_t)(1u << (i * CN + rn));
  const uint16_t release_mask = row_mask | col_mask;   // the CTAs that write into this CTA's ring (and vice versa)
  const int64_t cluster_id = blockIdx.x / CSIZE;
  const int64_t num_clusters = gridDim.x / CSIZE;

  if (tid == 0) {
    for (int s = 0; s < stages; s++) { oz_mbar_init(&full_bar[s], 1); oz_mbar_init(&empty_bar[s], CM + CN - 1); }
    oz_mbar_init(&acc_full, 1);
    oz_mbar_init(&acc_empty, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(oz_smem_u32(&tmem_base_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CSIZE > 1) oz_cluster_sync();   // every CTA's barriers exist before any peer multicasts into it
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  // Both single-thread roles run their loops in WARP-UNIFORM control flow and put only the asynchronous instructions under
  // one elected lane.  With the loops inside an `if (lane == 0)` branch ptxas cannot keep the descriptors in uniform
  // registers and wraps every UTCIMMA / UTMALDG in an ELECT + R2UR.BROADCAST + BRA.U.ANY loop (~150 clk of dependent
  // single-thread latency per instruction, plus a 64-bit division call per k-block for `it % stages`): the kernel was
  // bound by the issue rate of its MMA thread, at 41 % tensor-pipe activity, not by the pipe (0.50 clk per accumulator
  // column in tools/ubench_umma.cu) nor by TMA (profiles/r02_ozaki_issue_bound.md).
  if (warp == 0) {
    // ================= TMA producer =================
    const bool lead = oz_elect_one();
    int s = 0;            // ring slot
    uint32_t round = 0;   // number of times the ring has wrapped
    OzTile tl;
    uint64_t pol_keep;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
    for (int64_t t = cluster_id; oz_tile<OZ_BK, OZ_BN, CM, CN>(t, m_tiles, n_tiles, k_blocks, tri_mode, group_m, tl); t += num_clusters) {
      const int row0 = (int)((tl.mt * CM + rm) * OZ_BM), col0 = (tl.nt * CN + rn) * OZ_BN;
      for (int kb = tl.kb0; kb < tl.kb1; kb++) {
        if (round > 0) oz_mbar_wait(&empty_bar[s], (round - 1) & 1u);
        if (lead) {
          if (dbg & 2) {
            oz_mbar_arrive(&full_bar[s]);
          } else {
            oz_mbar_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
            uint8_t* st = smem + (size_t)s * stage_bytes;
            if (CSIZE == 1 && (dbg & 4)) {
              // the maps were encoded with a box depth of G slices: the slice-major stage layout is the same
              oz_tma_3d(st, &mapA, &full_bar[s], kb * OZ_BK, row0, 0);
              oz_tma_3d_hint(st + G * OZ_A_TILE, &mapB, &full_bar[s], kb * OZ_BK, col0, 0, pol_keep);
            } else if (CSIZE == 1 && l2_hints) {
              // the B slices (the triangular factor, re-used by every row tile of the launch) are kept in L2 with
              // evict_last; the A slices are only re-used by the column tiles of the current row group
              for (int p = 0; p < G; p++) oz_tma_3d(st + p * OZ_A_TILE, &mapA, &full_bar[s], kb * OZ_BK, row0, p);
              for (int q = 0; q < G; q++)
                oz_tma_3d_hint(st + G * OZ_A_TILE + q * OZ_B_TILE, &mapB, &full_bar[s], kb * OZ_BK, col0, q, pol_keep);
            } else if (CSIZE == 1) {
              for (int p = 0; p < G; p++) oz_tma_3d(st + p * OZ_A_TILE, &mapA, &full_bar[s], kb * OZ_BK, row0, p);
              for (int q = 0; q < G; q++) oz_tma_3d(st + G * OZ_A_TILE + q * OZ_B_TILE, &mapB, &full_bar[s], kb * OZ_BK, col0, q);
            } else {
              // each CTA fetches 1/CN of its row's A slices and 1/CM of its column's B slices and multicasts them
              for (int p = rn; p < G; p += CN) oz_tma_3d_mc(st + p * OZ_A_TILE, &mapA, &full_bar[s], kb * OZ_BK, row0, p, row_mask);
              for (int q = rm; q < G; q += CM) oz_tma_3d_mc(st + G * OZ_A_TILE + q * OZ_B_TILE, &mapB, &full_bar[s], kb * OZ_BK, col0, q, col_mask);
            }
          }
        }
        __syncwarp();
        if (++s == stages) { s = 0; round++; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const bool lead = oz_elect_one();
    // instruction descriptor: D = S32, A = B = signed int8, both K-major, M = 128; N is filled in per instruction
    const uint32_t idesc_base = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_BM >> 4) << 24);
    const uint64_t desc0 = oz_desc<OZ_BK>(smem);   // descriptor of ring slot 0; tiles advance its 16-byte address field
    int s = 0;
    uint32_t round = 0;
    int64_t tile_i = 0;
    OzTile tl;
    for (int64_t t = cluster_id; oz_tile<OZ_BK, OZ_BN, CM, CN>(t, m_tiles, n_tiles, k_blocks, tri_mode, group_m, tl); t += num_clusters, tile_i++) {
      if (tile_i > 0) {  // accumulators must have been drained by the epilogue of the previous tile
        oz_mbar_wait(&acc_empty, (uint32_t)((tile_i - 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      for (int kb = tl.kb0; kb < tl.kb1; kb++) {
        oz_mbar_wait(&full_bar[s], round & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lead) {
          const uint64_t ds = desc0 + (uint64_t)((uint32_t)(s * stage_bytes) >> 4);
          const bool first = (kb == tl.kb0);
          // A_p times the stacked tiles [B_0; ...; B_{G-1-p}] lands in the consecutive accumulators g = p .. G-1, so the
          // G(G+1)/2 slice products of a K=32 step are issued as wide UMMAs (N up to 256) instead of G(G+1)/2 N=64 ones:
          // the A tile is read from shared memory once per p, not once per pair.
          if (!(dbg & 1)) {
#pragma unroll
            for (int k = 0; k < OZ_BK / 32; k++) {
              for (int p = 0; p < G; p++) {
                const uint64_t ad = ds + (uint64_t)((uint32_t)(p * OZ_A_TILE) >> 4) + 2 * k;
                const int ncols = (G - p) * OZ_BN;
                for (int c0 = 0; c0 < ncols; c0 += 256) {
                  const int nn = (ncols - c0 < 256) ? (ncols - c0) : 256;
                  // stacked slices are contiguous rows behind the G A tiles
                  const uint64_t bd = ds + (uint64_t)((uint32_t)(G * OZ_A_TILE + c0 * OZ_BK) >> 4) + 2 * k;
                  const uint32_t idesc_n = idesc_base | ((uint32_t)(nn >> 3) << 17);
                  oz_umma(tmem_base + (uint32_t)(p * OZ_BN + c0), ad, bd, idesc_n, (first && p == 0 && k == 0) ? 0u : 1u);
                }
              }
            }
          }
          if (CSIZE == 1) oz_commit(&empty_bar[s]);
          else oz_commit_mc(&empty_bar[s], release_mask);
        }
        __syncwarp();
        if (++s == stages) { s = 0; round++; }
      }
      if (lead) oz_commit(&acc_full);
      __syncwarp();
    }
  } else {
    // ================= epilogue (warps 2..5) =================
    oz_epilogue<OZ_BK, OZ_BN, CM, CN>(mapC, tri_mode, M, N, G, row_scale, col_scale, C, ldc, l2_hints, group_m, tma_store, m_tiles,
                                      n_tiles, k_blocks, cluster_id, num_clusters, rm, rn, warp, lane, tid, tmem_base, acc_full,
                                      acc_empty, s_col, smem + (size_t)stages * stage_bytes);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CSIZE > 1) oz_cluster_sync();   // no CTA leaves while a peer may still signal its barriers
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// ---- split-ring variant (experiment, MCACQ_OZ_RING=1; measured SLOWER) ---------------------------------------------
// Same tile, same instructions, finer pipeline.  With whole k-blocks as ring stages only two 80 KB stages fit, so while the
// tensor pipe works on one stage exactly ONE refill is in flight: 9.1 ms at G = 6 with MMA-only 7.5 ms and TMA-only 5.7 ms
// (profiles/r02_ozaki_variants_uniform_issue.txt).  Here the B slices of a k-block (G x BN rows) are double-buffered as a
// block and the A slices (128 rows each) go through a ring of NA 8 KB slots with one mbarrier pair per slot, so that 2-3
// k-blocks of A tiles are in flight.  Result (profiles/r02_ozaki_ring.txt): 11.8 ms instead of 8.8 ms at G = 6 -- the G
// extra wait / fence / commit round trips per k-block stall the instruction stream of BOTH single-lane roles (MMA-only
// 11.0 ms, TMA-only 10.2 ms): barrier traffic costs more than the deeper pipeline gains.  Kept for the record.
constexpr int OZ_MAX_NA = 20;

template <int OZ_BN>
__global__ void __launch_bounds__(OZ_THREADS, 1)
ozaki_ring_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                  const __grid_constant__ CUtensorMap mapC, int tri_mode, int64_t M, int N, int K, int G, int NA,
                  const double* __restrict__ row_scale, const double* __restrict__ col_scale, double* __restrict__ C,
                  int64_t ldc, int l2_hints, int group_m, int dbg, int tma_store) {
  constexpr int OZ_BK = 64;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int OZ_A_TILE = OZ_BM * OZ_BK, OZ_B_TILE = OZ_BN * OZ_BK;
  const int b_block = G * OZ_B_TILE;            // one k-block of stacked B slices
  uint8_t* sB = smem;                           // [2][G][BN x 64 B]
  uint8_t* sA = smem + 2 * (size_t)b_block;     // [NA][128 x 64 B]
  __shared__ __align__(8) uint64_t fullA[OZ_MAX_NA], emptyA[OZ_MAX_NA], fullB[2], emptyB[2], acc_full, acc_empty;
  __shared__ uint32_t tmem_base_s;
  __shared__ double s_col[OZ_MAX_BN];
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int64_t m_tiles = (M + OZ_BM - 1) / OZ_BM;
  const int n_tiles = (N + OZ_BN - 1) / OZ_BN;
  const int k_blocks = (K + OZ_BK - 1) / OZ_BK;
  const int64_t cta_id = blockIdx.x, num_ctas = gridDim.x;

  if (tid == 0) {
    for (int i = 0; i < NA; i++) { oz_mbar_init(&fullA[i], 1); oz_mbar_init(&emptyA[i], 1); }
    for (int i = 0; i < 2; i++) { oz_mbar_init(&fullB[i], 1); oz_mbar_init(&emptyB[i], 1); }
    oz_mbar_init(&acc_full, 1);
    oz_mbar_init(&acc_empty, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(oz_smem_u32(&tmem_base_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ================= TMA producer (warp-uniform loops, one elected lane issues) =================
    const bool lead = oz_elect_one();
    int aslot = 0, bbuf = 0;
    uint32_t around = 0, bround = 0;   // wrap counts of the A ring / B double buffer
    OzTile tl;
    uint64_t pol_keep;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
    for (int64_t t = cta_id; oz_tile<OZ_BK, OZ_BN, 1, 1>(t, m_tiles, n_tiles, k_blocks, tri_mode, group_m, tl); t += num_ctas) {
      const int row0 = (int)(tl.mt * OZ_BM), col0 = tl.nt * OZ_BN;
      for (int kb = tl.kb0; kb < tl.kb1; kb++) {
        // B block of this k-block
        if (bround > 0) oz_mbar_wait(&emptyB[bbuf], (bround - 1) & 1u);
        if (lead) {
          if (dbg & 2) {
            oz_mbar_arrive(&fullB[bbuf]);
          } else {
            oz_mbar_expect_tx(&fullB[bbuf], (uint32_t)b_block);
            uint8_t* dst = sB + (size_t)bbuf * b_block;
            // the B slices (the triangular factor, re-used by every row tile of the launch) stay in L2 with evict_last
            for (int q = 0; q < G; q++) {
              if (l2_hints) oz_tma_3d_hint(dst + q * OZ_B_TILE, &mapB, &fullB[bbuf], kb * OZ_BK, col0, q, pol_keep);
              else oz_tma_3d(dst + q * OZ_B_TILE, &mapB, &fullB[bbuf], kb * OZ_BK, col0, q);
            }
          }
        }
        __syncwarp();
        bbuf ^= 1;
        if (bbuf == 0) bround++;
        // A slices, one ring slot each
        for (int p = 0; p < G; p++) {
          if (around > 0) oz_mbar_wait(&emptyA[aslot], (around - 1) & 1u);
          if (lead) {
            if (dbg & 2) {
              oz_mbar_arrive(&fullA[aslot]);
            } else {
              oz_mbar_expect_tx(&fullA[aslot], (uint32_t)OZ_A_TILE);
              oz_tma_3d(sA + (size_t)aslot * OZ_A_TILE, &mapA, &fullA[aslot], kb * OZ_BK, row0, p);
            }
          }
          __syncwarp();
          if (++aslot == NA) { aslot = 0; around++; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const bool lead = oz_elect_one();
    const uint32_t idesc_base = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_BM >> 4) << 24);
    const uint64_t descB0 = oz_desc<OZ_BK>(sB), descA0 = oz_desc<OZ_BK>(sA);
    int aslot = 0, bbuf = 0;
    uint32_t around = 0, bround = 0;
    int64_t tile_i = 0;
    OzTile tl;
    for (int64_t t = cta_id; oz_tile<OZ_BK, OZ_BN, 1, 1>(t, m_tiles, n_tiles, k_blocks, tri_mode, group_m, tl); t += num_ctas, tile_i++) {
      if (tile_i > 0) {  // accumulators must have been drained by the epilogue of the previous tile
        oz_mbar_wait(&acc_empty, (uint32_t)((tile_i - 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      for (int kb = tl.kb0; kb < tl.kb1; kb++) {
        oz_mbar_wait(&fullB[bbuf], bround & 1u);
        const uint64_t db = descB0 + (uint64_t)((uint32_t)(bbuf * b_block) >> 4);
        const bool first = (kb == tl.kb0);
        for (int p = 0; p < G; p++) {
          oz_mbar_wait(&fullA[aslot], around & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (lead) {
            // A_p times the stacked tiles [B_0; ...; B_{G-1-p}] lands in the consecutive accumulators g = p .. G-1: wide
            // UMMAs (N up to 256); the very first instruction of a tile (p = 0, k = 0) covers all G * BN columns and
            // overwrites them, everything after accumulates
            if (!(dbg & 1)) {
              const uint64_t da = descA0 + (uint64_t)((uint32_t)(aslot * OZ_A_TILE) >> 4);
              const int ncols = (G - p) * OZ_BN;
#pragma unroll
              for (int k = 0; k < OZ_BK / 32; k++) {
                for (int c0 = 0; c0 < ncols; c0 += 256) {
                  const int nn = (ncols - c0 < 256) ? (ncols - c0) : 256;
                  const uint64_t bd = db + (uint64_t)((uint32_t)(c0 * OZ_BK) >> 4) + 2 * k;   // stacked slices are contiguous rows
                  const uint32_t idesc_n = idesc_base | ((uint32_t)(nn >> 3) << 17);
                  oz_umma(tmem_base + (uint32_t)(p * OZ_BN + c0), da + 2 * k, bd, idesc_n, (first && p == 0 && k == 0) ? 0u : 1u);
                }
              }
            }
            oz_commit(&emptyA[aslot]);
          }
          __syncwarp();
          if (++aslot == NA) { aslot = 0; around++; }
        }
        if (lead) oz_commit(&emptyB[bbuf]);
        __syncwarp();
        bbuf ^= 1;
        if (bbuf == 0) bround++;
      }
      if (lead) oz_commit(&acc_full);
      __syncwarp();
    }
  } else {
    oz_epilogue<OZ_BK, OZ_BN, 1, 1>(mapC, tri_mode, M, N, G, row_scale, col_scale, C, ldc, l2_hints, group_m, tma_store, m_tiles,
                                    n_tiles, k_blocks, cta_id, num_ctas, 0, 0, warp, lane, tid, tmem_base, acc_full, acc_empty, s_col,
                                    sA + (size_t)NA * OZ_A_TILE);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// ---- CTA-pair variant (cta_group::2) --------------------------------------------------------------------------------
// Two CTAs of a cluster (one TPC) work on a 256-row x BN tile: each loads its own 128 rows of the A slices but only HALF of
// every B slice tile (BN/2 rows), and the leader's `tcgen05.mma.cta_group::2` (M = 256) reads both halves, so the operand
// bytes an SM has to ingest per output drop from G(128 + BN) to G(128 + BN/2).  The N halves of the two CTAs interleave
// inside one instruction, so slices cannot be stacked into wide UMMAs here: one M=256 x N=BN instruction per slice pair.
// Barrier plumbing as in CUTLASS's 2-SM pipelines: both producers signal the LEADER's full barrier (armed by the leader
// with the bytes of both CTAs), the leader's commits are multicast to both CTAs' empty / accumulator barriers, and both
// CTAs' epilogue warps arrive on the leader's accumulator-empty barrier.
__device__ __forceinline__ uint32_t oz_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void oz_tma_3d_2sm(void* dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(oz_smem_u32(dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void oz_umma2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void oz_commit2_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(oz_smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void oz_mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

template <int OZ_BK, int OZ_BN>
__global__ void __launch_bounds__(OZ_THREADS, 1)
ozaki_imma2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapBh, int tri_mode,
                   int64_t M, int N, int K, int G, int stages, const double* __restrict__ row_scale,
                   const double* __restrict__ col_scale, double* __restrict__ C, int64_t ldc, int group_m) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int OZ_A_TILE = OZ_BM * OZ_BK, OZ_BH_TILE = (OZ_BN / 2) * OZ_BK;
  const int stage_bytes = G * (OZ_A_TILE + OZ_BH_TILE);
  __shared__ __align__(8) uint64_t full_bar[OZ_MAX_STAGES], empty_bar[OZ_MAX_STAGES], acc_full, acc_empty;
  __shared__ uint32_t tmem_base_s;
  __shared__ double s_col[OZ_MAX_BN];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m_tiles = (M + OZ_BM - 1) / OZ_BM;
  const int n_tiles = (N + OZ_BN - 1) / OZ_BN;
  const int k_blocks = (K + OZ_BK - 1) / OZ_BK;
  const uint32_t crank = oz_cluster_rank();        // 0 = leader (issues the MMAs), 1 = peer
  const int64_t cluster_id = blockIdx.x / 2;
  const int64_t num_clusters = gridDim.x / 2;

  if (tid == 0) {
    for (int s = 0; s < stages; s++) { oz_mbar_init(&full_bar[s], 1); oz_mbar_init(&empty_bar[s], 1); }
    oz_mbar_init(&acc_full, 1);
    oz_mbar_init(&acc_empty, 8);   // four epilogue warps in each CTA of the pair (only the leader's copy is used)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(oz_smem_u32(&tmem_base_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  oz_cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      // ================= TMA producer (both CTAs) =================
      int64_t it = 0;
      OzTile tl;
      for (int64_t t = cluster_id; oz_tile<OZ_BK, OZ_BN, 2, 1>(t, m_tiles, n_tiles, k_blocks, tri_mode, group_m, tl); t += num_clusters) {
        const int row0 = (int)((tl.mt * 2 + crank) * OZ_BM);
        const int colh0 = tl.nt * OZ_BN + (int)crank * (OZ_BN / 2);
        for (int kb = tl.kb0; kb < tl.kb1; kb++, it++) {
          const int s = (int)(it % stages);
          if (it >= stages) oz_mbar_wait(&empty_bar[s], (uint32_t)(((it / stages) - 1) & 1));
          if (crank == 0) oz_mbar_expect_tx(&full_bar[s], (uint32_t)(2 * stage_bytes));
          const uint32_t lbar = oz_mapa(oz_smem_u32(&full_bar[s]), 0);
          uint8_t* st = smem + (size_t)s * stage_bytes;
          for (int p = 0; p < G; p++) oz_tma_3d_2sm(st + p * OZ_A_TILE, &mapA, lbar, kb * OZ_BK, row0, p);
          for (int q = 0; q < G; q++) oz_tma_3d_2sm(st + G * OZ_A_TILE + q * OZ_BH_TILE, &mapBh, lbar, kb * OZ_BK, colh0, q);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && crank == 0) {
      // ================= MMA issuer (leader CTA only) =================
      // D = S32, A = B = signed int8, both K-major, M = 256 over the CTA pair, N = BN
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      int64_t it = 0, tile_i = 0;
      OzTile tl;
      for (int64_t t = cluster_id; oz_tile<OZ_BK, OZ_BN, 2, 1>(t, m_tiles, n_tiles, k_blocks, tri_mode, group_m, tl); t += num_clusters, tile_i++) {
        if (tile_i > 0) {
          oz_mbar_wait(&acc_empty, (uint32_t)((tile_i - 1) & 1));
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        for (int kb = tl.kb0; kb < tl.kb1; kb++, it++) {
          const int s = (int)(it % stages);
          oz_mbar_wait(&full_bar[s], (uint32_t)((it / stages) & 1));
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint8_t* st = smem + (size_t)s * stage_bytes;
          const bool first = (kb == tl.kb0);
#pragma unroll
          for (int k = 0; k < OZ_BK / 32; k++) {
            for (int p = 0; p < G; p++) {
              const uint64_t ad = oz_desc<OZ_BK>(st + p * OZ_A_TILE) + 2 * k;
              for (int q = 0; q + p < G; q++) {
                const uint64_t bd = oz_desc<OZ_BK>(st + G * OZ_A_TILE + q * OZ_BH_TILE) + 2 * k;
                oz_umma2(tmem_base + (uint32_t)((p + q) * OZ_BN), ad, bd, idesc, (first && p == 0 && k == 0) ? 0u : 1u);
              }
            }
          }
          oz_commit2_mc(&empty_bar[s], (uint16_t)3);
        }
        oz_commit2_mc(&acc_full, (uint16_t)3);
      }
    }
  } else {
    // ================= epilogue (warps 2..5, both CTAs: each drains its own 128 TMEM lanes) =================
    const int lg = warp & 3;
    const int r_in_tile = lg * 32 + lane;
    const int etid = tid - 64;
    const uint32_t leader_acc_empty = oz_mapa(oz_smem_u32(&acc_empty), 0);
    int64_t tile_i = 0;
    OzTile tl;
    for (int64_t t = cluster_id; oz_tile<OZ_BK, OZ_BN, 2, 1>(t, m_tiles, n_tiles, k_blocks, tri_mode, group_m, tl); t += num_clusters, tile_i++) {
      const int64_t row = (tl.mt * 2 + crank) * OZ_BM + r_in_tile;
      const int col0 = tl.nt * OZ_BN;
      for (int c = etid; c < OZ_BN; c += 128) s_col[c] = (col0 + c < N) ? col_scale[col0 + c] : 0.0;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const bool have_acc = tl.kb1 > tl.kb0;
      if (have_acc) {
        oz_mbar_wait(&acc_full, (uint32_t)(tile_i & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      const bool row_ok = row < M;
      const double rs = row_ok ? row_scale[row] : 0.0;
      double* dst = C + (row_ok ? row : 0) * ldc + col0;
      const bool vec_ok = (ldc & 1) == 0;
#pragma unroll 1
      for (int c0 = 0; c0 < OZ_BN; c0 += 16) {
        double acc[16];
#pragma unroll
        for (int j = 0; j < 16; j++) acc[j] = 0.0;
        if (have_acc) {
          uint32_t v[OZ_MAXG][16];
#pragma unroll
          for (int g = 0; g < OZ_MAXG; g++) {
            if (g < G) {
              const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(g * OZ_BN + c0);
              asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                           : "=r"(v[g][0]), "=r"(v[g][1]), "=r"(v[g][2]), "=r"(v[g][3]), "=r"(v[g][4]), "=r"(v[g][5]), "=r"(v[g][6]),
                             "=r"(v[g][7]), "=r"(v[g][8]), "=r"(v[g][9]), "=r"(v[g][10]), "=r"(v[g][11]), "=r"(v[g][12]),
                             "=r"(v[g][13]), "=r"(v[g][14]), "=r"(v[g][15])
                           : "r"(taddr));
            }
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          double w = 1.0 / 4294967296.0;
#pragma unroll
          for (int g0 = 0; g0 < OZ_MAXG; g0 += 3) {
            if (g0 < G) {
#pragma unroll
              for (int j = 0; j < 16; j++) {
                const long long s0 = (long long)(int32_t)v[g0][j];
                const long long s1 = (g0 + 1 < OZ_MAXG && g0 + 1 < G) ? (long long)(int32_t)v[g0 + 1 < OZ_MAXG ? g0 + 1 : g0][j] : 0ll;
                const long long s2 = (g0 + 2 < OZ_MAXG && g0 + 2 < G) ? (long long)(int32_t)v[g0 + 2 < OZ_MAXG ? g0 + 2 : g0][j] : 0ll;
                const long long tt = s0 * 65536ll + s1 * 256ll + s2;
                const double dd = __longlong_as_double(0x4338000000000000ll + tt) - 6755399441055744.0;
                acc[j] = fma(w, dd, acc[j]);
              }
              w *= (1.0 / 16777216.0);
            }
          }
        }
        if (row_ok && col0 + c0 < N) {
          if (col0 + c0 + 16 <= N && vec_ok) {
#pragma unroll
            for (int j = 0; j < 16; j += 2)
              __stcs(reinterpret_cast<double2*>(dst + c0 + j),
                     make_double2(acc[j] * rs * s_col[c0 + j], acc[j + 1] * rs * s_col[c0 + j + 1]));
          } else {
            for (int j = 0; j < 16; j++) if (col0 + c0 + j < N) dst[c0 + j] = acc[j] * rs * s_col[c0 + j];
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) oz_mbar_arrive_remote(leader_acc_empty);
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  oz_cluster_sync();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// ---- slicing ------------------------------------------------------------------------------------------------------
// X[rows x K] (fp64, pitch ldx) -> G int8 slices [G][rows][Kp] (Kp = slice pitch >= K, zero padded) and the row scale
// 2^(e+2) with |x| < 2^e.  fixed_exp != INT_MIN: use that exponent for every row (K(X, X_train) <= outputscale).
__global__ void __launch_bounds__(256)
slice_rows_kernel(const double* __restrict__ X, int64_t rows, int K, int64_t ldx, int Kp, int G, int fixed_exp,
                  int8_t* __restrict__ S, double* __restrict__ scale) {
  const int64_t row = blockIdx.x;
  if (row >= rows) return;
  const double* x = X + row * ldx;
  __shared__ double s_red[8];
  __shared__ int s_exp;
  int e;
  if (fixed_exp != INT_MIN) {
    e = fixed_exp;
  } else {
    double mx = 0.0;
    for (int k = threadIdx.x; k < K; k += 256) mx = fmax(mx, fabs(x[k]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
      double m = 0.0;
      for (int w = 0; w < 8; w++) m = fmax(m, s_red[w]);
      int ex = 0;
      if (m > 0.0 && isfinite(m)) { frexp(m, &ex); }  // m = f * 2^ex, f in [0.5, 1)  =>  m * 2^-ex < 1
      s_exp = ex;
    }
    __syncthreads();
    e = s_exp;
  }
  // x = 2^(e+2) * sum_p D_p 256^-(p+1) with balanced digits D_p in [-128, 127]: round x to a (8G-2)-bit fixed-point
  // integer relative to 2^e and peel signed bytes from the least significant end (carry-propagating, exact).  The
  // extra headroom bit keeps the carries inside G bytes: G balanced bytes only reach 0.498 * 256^G.
  if (threadIdx.x == 0) scale[row] = ldexp(1.0, e + 2);
  const size_t slice_stride = (size_t)rows * Kp;
  int8_t* out = S + row * (size_t)Kp;
  const int shift = 8 * G - 2 - e;
  const long long lim = (1ll << (8 * G - 2));  // |x| < 2^e  =>  |X| <= 2^(8G-2)
  for (int k0 = threadIdx.x * 4; k0 < Kp; k0 += 1024) {  // Kp is a multiple of 16: 4 consecutive columns per thread
    unsigned long long Y[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      long long X = (k0 + j < K) ? __double2ll_rn(ldexp(x[k0 + j], shift)) : 0ll;
      X = X > lim ? lim : (X < -lim ? -lim : X);
      Y[j] = balanced_bytes(X);
    }
    digits4(Y, [&](int dg, unsigned w) {
      if (dg < G) *reinterpret_cast<unsigned*>(out + (size_t)(G - 1 - dg) * slice_stride + k0) = w;
    });
  }
}

typedef CUresult (*OzEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static OzEncodeFn oz_encode_fn() {
  static OzEncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess) fn = (OzEncodeFn)p;
  }
  return fn;
}

static int oz_make_map(CUtensorMap* m, const void* ptr, uint64_t slices, uint64_t rows, uint64_t kbytes, uint64_t pitch,
                       uint32_t box_rows, int OZ_BK, uint32_t box_slices = 1) {
  OzEncodeFn enc = oz_encode_fn();
  if (!enc) return MCACQ_EINVAL;
  cuuint64_t dims[3] = {kbytes, rows, slices};
  cuuint64_t strides[2] = {pitch, pitch * rows};
  cuuint32_t box[3] = {(cuuint32_t)OZ_BK, box_rows, box_slices};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, OZ_BK == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : MCACQ_EINVAL;
}

static int oz_make_map_c(CUtensorMap* m, const double* C, uint64_t rows, uint64_t cols, uint64_t ldc) {
  OzEncodeFn enc = oz_encode_fn();
  if (!enc) return MCACQ_EINVAL;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ldc * 8};
  cuuint32_t box[2] = {16, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(C), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : MCACQ_EINVAL;
}

template <int BKB, int OZ_BN, int CM, int CN>
static int oz_launch_t(const CUtensorMap& mapA, const CUtensorMap& mapB, const CUtensorMap& mapC, int tma_store, int tri_mode,
                       int64_t M, int N, int K, int G,
                       int stages, size_t smem, size_t smem_budget, const double* row_scale, const double* col_scale,
                       double* C, int64_t ldc, cudaStream_t st, int dbg) {
  auto kern = ozaki_imma_kernel<BKB, OZ_BN, CM, CN>;
  static int max_clusters = -1;
  constexpr int CS = CM * CN;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  cfg.blockDim = dim3(OZ_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (CS > 1) ? 1 : 0;
  if (max_clusters < 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem_budget + 1024));
    if (e != cudaSuccess) return (int)e;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (CS > 1) {
      cfg.gridDim = dim3(sms / CS * CS);
      int n = 0;
      e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
      if (e != cudaSuccess || n <= 0) { cudaGetLastError(); return e != cudaSuccess ? (int)e : MCACQ_ELIMIT; }
      max_clusters = n;
    } else {
      max_clusters = sms;
    }
  }
  const int l2_hints = (getenv("MCACQ_OZ_HINT") != nullptr) ? atoi(getenv("MCACQ_OZ_HINT")) : 1;
  const int group_m = (getenv("MCACQ_OZ_GROUP") != nullptr) ? atoi(getenv("MCACQ_OZ_GROUP")) : OZ_GROUP_M;
  const int64_t m_tiles = (M + OZ_BM - 1) / OZ_BM;
  const int64_t n_tiles = (N + OZ_BN - 1) / OZ_BN;
  const int64_t ctiles = ((m_tiles + CM - 1) / CM) * ((n_tiles + CN - 1) / CN);
  const int nclusters = (int)(ctiles < max_clusters ? ctiles : max_clusters);
  cfg.gridDim = dim3(nclusters * CS);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, mapA, mapB, mapC, tri_mode, M, N, K, G, stages, row_scale, col_scale, C, ldc, l2_hints, group_m, dbg,
                                     tma_store);
  count_launch();
  if (e != cudaSuccess) return (int)e;
  MCACQ_CUDA_CHECK_LAUNCH();
  return 0;
}

template <int BN>
static int oz_launch2_t(const CUtensorMap& mapA, const CUtensorMap& mapBh, int tri_mode, int64_t M, int N, int K, int G,
                        int stages, size_t smem, size_t smem_budget, const double* row_scale, const double* col_scale, double* C,
                        int64_t ldc, cudaStream_t st) {
  auto kern = ozaki_imma2_kernel<64, BN>;
  static int max_clusters = -1;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  cfg.blockDim = dim3(OZ_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (max_clusters < 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem_budget + 1024));
    if (e != cudaSuccess) return (int)e;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cfg.gridDim = dim3(sms / 2 * 2);
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    if (e != cudaSuccess || n <= 0) { cudaGetLastError(); return e != cudaSuccess ? (int)e : MCACQ_ELIMIT; }
    max_clusters = n;
  }
  const int group_m = (getenv("MCACQ_OZ_GROUP") != nullptr) ? atoi(getenv("MCACQ_OZ_GROUP")) : OZ_GROUP_M;
  const int64_t m_tiles = (M + OZ_BM - 1) / OZ_BM;
  const int64_t n_tiles = (N + BN - 1) / BN;
  const int64_t ctiles = ((m_tiles + 1) / 2) * n_tiles;
  const int nclusters = (int)(ctiles < max_clusters ? ctiles : max_clusters);
  cfg.gridDim = dim3(nclusters * 2);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, mapA, mapBh, tri_mode, M, N, K, G, stages, row_scale, col_scale, C, ldc, group_m);
  count_launch();
  if (e != cudaSuccess) return (int)e;
  MCACQ_CUDA_CHECK_LAUNCH();
  return 0;
}

template <int OZ_BN>
static int oz_launch_ring_t(const CUtensorMap& mapA, const CUtensorMap& mapB, const CUtensorMap& mapC, int tma_store, int tri_mode,
                            int64_t M, int N, int K, int G, int NA, size_t smem, size_t smem_budget, const double* row_scale,
                            const double* col_scale, double* C, int64_t ldc, cudaStream_t st, int dbg) {
  auto kern = ozaki_ring_kernel<OZ_BN>;
  static int sms = -1;
  if (sms < 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem_budget + 1024));
    if (e != cudaSuccess) return (int)e;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int l2_hints = (getenv("MCACQ_OZ_HINT") != nullptr) ? atoi(getenv("MCACQ_OZ_HINT")) : 1;
  const int group_m = (getenv("MCACQ_OZ_GROUP") != nullptr) ? atoi(getenv("MCACQ_OZ_GROUP")) : OZ_GROUP_M;
  const int64_t tiles = ((M + OZ_BM - 1) / OZ_BM) * ((N + OZ_BN - 1) / OZ_BN);
  const int grid = (int)(tiles < sms ? tiles : sms);
  kern<<<grid, OZ_THREADS, smem, st>>>(mapA, mapB, mapC, tri_mode, M, N, K, G, NA, row_scale, col_scale, C, ldc, l2_hints, group_m,
                                       dbg, tma_store);
  count_launch();
  MCACQ_CUDA_CHECK_LAUNCH();
  return 0;
}

static int oz_launch(int bk, int bn, int cm, int cn, const CUtensorMap& mapA, const CUtensorMap& mapB, const CUtensorMap& mapC,
                     int tma_store, int tri_mode, int64_t M,
                     int N, int K, int G, int stages, size_t smem, size_t smem_budget, const double* row_scale,
                     const double* col_scale, double* C, int64_t ldc, cudaStream_t st, int dbg) {
#define OZ_CASE(B, W, A_, C_) if (bk == B && bn == W && cm == A_ && cn == C_) \
    return oz_launch_t<B, W, A_, C_>(mapA, mapB, mapC, tma_store, tri_mode, M, N, K, G, stages, smem, smem_budget, row_scale, col_scale, C, ldc, st, dbg);
  OZ_CASE(64, 64, 1, 1) OZ_CASE(64, 80, 1, 1) OZ_CASE(64, 96, 1, 1) OZ_CASE(64, 128, 1, 1) OZ_CASE(64, 160, 1, 1) OZ_CASE(64, 256, 1, 1)
  OZ_CASE(128, 64, 1, 1) OZ_CASE(128, 128, 1, 1) OZ_CASE(128, 160, 1, 1) OZ_CASE(128, 256, 1, 1)
  // cluster (TMA multicast) variants, kept for the record: measured equal or slower (profiles/r01_ozaki_int8.md)
  OZ_CASE(64, 64, 2, 2) OZ_CASE(64, 64, 1, 4) OZ_CASE(64, 64, 2, 1) OZ_CASE(64, 64, 1, 2)
#undef OZ_CASE
  return MCACQ_EINVAL;
}

// Column-tile width: the widest multiple of 16 whose G int32 accumulator tiles fit the 512 TMEM columns (<= 256, the
// UMMA / TMA box limit).
static int oz_pick_bn(int G) {
  int bn = (512 / G) / 16 * 16;
  return bn > OZ_MAX_BN ? OZ_MAX_BN : bn;
}

}  // namespace mcacq

extern "C" int mcacq_slice_rows(const double* X, int64_t rows, int K, int64_t ldx, int Kp, int G, int use_fixed_exp,
                                int fixed_exp, int8_t* slices, double* row_scale, void* stream) {
  using namespace mcacq;
  if (!X || !slices || !row_scale || rows < 0 || K <= 0 || Kp < K || ldx < K || G <= 0 || G > OZ_MAXG || (Kp % 4) != 0)
    return MCACQ_EINVAL;
  if (rows == 0) return 0;
  slice_rows_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(X, rows, K, ldx, Kp, G,
                                                                       use_fixed_exp ? fixed_exp : INT_MIN, slices, row_scale);
  count_launch();
  MCACQ_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" int mcacq_ozaki_contract(int tri_mode, int64_t M, int N, int K, int G, const int8_t* A_slices,
                                    const double* row_scale, const int8_t* B_slices, const double* col_scale, double* C,
                                    int64_t ldc, void* stream) {
  using namespace mcacq;
  if (!A_slices || !B_slices || !row_scale || !col_scale || !C || M < 0 || N <= 0 || K <= 0 || ldc < N) return MCACQ_EINVAL;
  if (G <= 0 || G > OZ_MAXG || (K % 16) != 0 || tri_mode < 0 || tri_mode > 2) return MCACQ_EINVAL;
  if ((int64_t)K * 128 * 128 * G >= (int64_t)1 << 31) return MCACQ_ELIMIT;  // exact int32 accumulation of a diagonal
  if (M == 0) return 0;
  // 128-byte k-blocks (full L2 lines, SWIZZLE_128B) when at least two stages of them fit, else 64-byte k-blocks
  const size_t smem_budget = 232448 - 4096 - 1024;  // 227 KB per CTA minus static shared memory (column scales, barriers) and alignment slack
  int bn = (getenv("MCACQ_OZ_BN") != nullptr) ? atoi(getenv("MCACQ_OZ_BN")) : oz_pick_bn(G);
  if (bn < 16 || bn > OZ_MAX_BN || (bn % 16) != 0 || G * bn > 512) return MCACQ_EINVAL;
  int bk = (getenv("MCACQ_OZ_BK") != nullptr) ? atoi(getenv("MCACQ_OZ_BK")) : 0;
  // the epilogue stages its output for TMA stores: 4 warps x 2 buffers x 4 KB (+ 1 KB alignment) behind the operand ring
  const size_t epi_bytes = 4 * 2 * 4096 + 1024;
  if (bk != 64 && bk != 128) bk = (bn != 80 && bn != 96 && (size_t)2 * G * (OZ_BM + bn) * 128 <= smem_budget - epi_bytes) ? 128 : 64;
  const size_t stage_bytes = (size_t)G * (OZ_BM + bn) * bk;
  int stages = (int)((smem_budget - epi_bytes) / stage_bytes);
  if (stages > OZ_MAX_STAGES) stages = OZ_MAX_STAGES;
  if (getenv("MCACQ_OZ_STAGES") != nullptr) { int v = atoi(getenv("MCACQ_OZ_STAGES")); if (v >= 1 && v < stages) stages = v; }
  if (stages < 1) return MCACQ_ELIMIT;
  const size_t smem = (size_t)stages * stage_bytes + 1024 + epi_bytes;
  const int cta2 = (getenv("MCACQ_OZ_CTA2") != nullptr) ? atoi(getenv("MCACQ_OZ_CTA2")) : 0;
  if (cta2 && (bn == 64 || bn == 80 || bn == 96 || bn == 128)) {
    // CTA-pair variant: 64-byte k-blocks, half B tiles per CTA
    const size_t sb = (size_t)G * (OZ_BM + bn / 2) * 64;
    int st2 = (int)(smem_budget / sb);
    if (st2 > OZ_MAX_STAGES) st2 = OZ_MAX_STAGES;
    if (st2 < 1) return MCACQ_ELIMIT;
    CUtensorMap mA, mBh;
    int rc2;
    if ((rc2 = oz_make_map(&mA, A_slices, (uint64_t)G, (uint64_t)M, (uint64_t)K, (uint64_t)K, OZ_BM, 64))) return rc2;
    if ((rc2 = oz_make_map(&mBh, B_slices, (uint64_t)G, (uint64_t)N, (uint64_t)K, (uint64_t)K, (uint32_t)(bn / 2), 64))) return rc2;
    const size_t sm2 = (size_t)st2 * sb + 1024;
    cudaStream_t s2 = (cudaStream_t)stream;
    if (bn == 64) return oz_launch2_t<64>(mA, mBh, tri_mode, M, N, K, G, st2, sm2, smem_budget, row_scale, col_scale, C, ldc, s2);
    if (bn == 80) return oz_launch2_t<80>(mA, mBh, tri_mode, M, N, K, G, st2, sm2, smem_budget, row_scale, col_scale, C, ldc, s2);
    if (bn == 96) return oz_launch2_t<96>(mA, mBh, tri_mode, M, N, K, G, st2, sm2, smem_budget, row_scale, col_scale, C, ldc, s2);
    return oz_launch2_t<128>(mA, mBh, tri_mode, M, N, K, G, st2, sm2, smem_budget, row_scale, col_scale, C, ldc, s2);
  }
  CUtensorMap mapA, mapB;
  int rc;
  int dbg = (getenv("MCACQ_OZ_DEBUG") != nullptr) ? atoi(getenv("MCACQ_OZ_DEBUG")) : 0;
  // One TMA box per operand and k-block spanning all G slices (the slice-major stage layout is exactly a 3-D box): 2
  // instead of 2 G bulk-tensor instructions per k-block, 8-13 % faster (G = 6: 9.15 -> 8.29 ms, G = 7: 12.1 -> 10.5 ms).
  // Kernel flag bit 2; MCACQ_OZ_BOX=0 restores one box per slice.  Cluster (multicast) variants split the slices between
  // the CTAs and keep per-slice boxes.
  int cm_env = 1, cn_env = 1;
  if (getenv("MCACQ_OZ_CLUSTER") != nullptr) { int v = atoi(getenv("MCACQ_OZ_CLUSTER")); cm_env = v / 10; cn_env = v % 10; }
  const int one_box = (getenv("MCACQ_OZ_BOX") != nullptr) ? atoi(getenv("MCACQ_OZ_BOX")) : 1;
  const int use_ring = (getenv("MCACQ_OZ_RING") != nullptr) ? atoi(getenv("MCACQ_OZ_RING")) : 0;
  if (one_box && cm_env * cn_env == 1 && !use_ring) dbg |= 4; else dbg &= ~4;
  const uint32_t box_slices = (dbg & 4) ? (uint32_t)G : 1u;
  if ((rc = oz_make_map(&mapA, A_slices, (uint64_t)G, (uint64_t)M, (uint64_t)K, (uint64_t)K, OZ_BM, bk, box_slices))) return rc;
  if ((rc = oz_make_map(&mapB, B_slices, (uint64_t)G, (uint64_t)N, (uint64_t)K, (uint64_t)K, (uint32_t)bn, bk, box_slices))) return rc;
  // Cluster shape (rows x columns of CTA tiles sharing operand loads through TMA multicast).  Measured on B200
  // (profiles/r01_ozaki_int8.md): multicast halves the L2 reads but not the bytes delivered to each SM, which is what
  // bounds this kernel (~20 B/clk/SM, ~5.8 TB/s chip-wide), so 1x1 is the default; 2x1 / 1x2 tie, 2x2 is 13% slower.
  int cm = 1, cn = 1;
  if (getenv("MCACQ_OZ_CLUSTER") != nullptr) { int v = atoi(getenv("MCACQ_OZ_CLUSTER")); cm = v / 10; cn = v % 10; }
  if (cm * cn > 1 && (bn != 64 || bk != 64)) return MCACQ_EINVAL;
  if (cm * cn > 1 && (dbg & 4)) return MCACQ_EINVAL;
  // output map (fp64, 16 x 32 boxes, SWIZZLE_128B); TMA needs a 16-byte aligned base and row pitch, otherwise the
  // epilogue falls back to per-thread stores
  CUtensorMap mapC = mapA;
  int tma_store = ((ldc & 1) == 0 && ((uintptr_t)C & 15) == 0) ? 1 : 0;
  if (getenv("MCACQ_OZ_TMASTORE") != nullptr) tma_store = tma_store && atoi(getenv("MCACQ_OZ_TMASTORE"));
  if (tma_store && oz_make_map_c(&mapC, C, (uint64_t)M, (uint64_t)N, (uint64_t)ldc) != 0) tma_store = 0;
  // split-ring pipeline (64-byte k-blocks, no clusters): experiment, off by default (slower, see the kernel's comment)
  if (use_ring && bk == 64 && cm * cn == 1 && (bn == 64 || bn == 80 || bn == 96 || bn == 128)) {
    const size_t b_bytes = (size_t)2 * G * bn * 64;
    int NA = (int)((smem_budget - epi_bytes - b_bytes) / (OZ_BM * 64));
    if (NA > OZ_MAX_NA) NA = OZ_MAX_NA;
    if (getenv("MCACQ_OZ_NA") != nullptr) { int v = atoi(getenv("MCACQ_OZ_NA")); if (v >= 1 && v < NA) NA = v; }
    if (NA >= 2) {
      const size_t smem_ring = b_bytes + (size_t)NA * OZ_BM * 64 + 1024 + epi_bytes;
      cudaStream_t sr = (cudaStream_t)stream;
#define OZ_RING_CASE(W) if (bn == W) return oz_launch_ring_t<W>(mapA, mapB, mapC, tma_store, tri_mode, M, N, K, G, NA, smem_ring, \
                                                                smem_budget, row_scale, col_scale, C, ldc, sr, dbg);
      OZ_RING_CASE(64) OZ_RING_CASE(80) OZ_RING_CASE(96) OZ_RING_CASE(128)
#undef OZ_RING_CASE
    }
  }
  return oz_launch(bk, bn, cm, cn, mapA, mapB, mapC, tma_store, tri_mode, M, N, K, G, stages, smem, smem_budget, row_scale, col_scale, C, ldc,
                   (cudaStream_t)stream, dbg);
}
