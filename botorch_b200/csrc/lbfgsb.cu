// Device-resident batched L-BFGS-B: one CTA per restart, reverse communication through device buffers.
//
// Replaces the host loop of the reference's batched optimiser -- botorch/optim/batched_lbfgs_b.py:365-634 steps one scipy
// `setulb` state machine per restart on the HOST and evaluates all active restarts together, so every round moves
// b' x q x d doubles host->device and b' (1 + q d) doubles back (botorch/generation/gen.py:423-485).  Here the state
// machines live in HBM: `mcacq_lbfgsb_step` consumes (acq, d acq / dX) straight from the fused forward / backward kernels
// and writes the next trial points into the same X buffer, so an optimiser round is  forward -> backward -> step  with no
// host hop; the host only polls the number of active restarts every few rounds.
//
// Algorithm: L-BFGS-B (Byrd, Lu, Nocedal, Zhu 1995; Morales, Nocedal 2011) exactly as restated in oracle/lbfgsb.py -- the
// operation order below follows that file function by function (cauchy, subsm, dcsrch / dcstep, the update of the compact
// matrices, the stopping tests), which in turn tracks scipy's iterates to rounding (tests/test_lbfgsb_device_model.py).
// Deviations from scipy are reduction orders only (warp shuffle trees instead of sequential sums).
//
// Work split inside a CTA (128 threads): D-length vector work is strided over the threads with fixed-order warp / CTA
// reductions; the 2m x 2m (<= 20 x 20) dense algebra and the scalar line-search state machine run on thread 0 with the
// operands in shared memory.
#include <cfloat>
#include <cmath>
#include <math_constants.h>
#include "common.cuh"

namespace mcacq {

constexpr int LB_THREADS = 128;
constexpr int LB_WARPS = LB_THREADS / 32;
constexpr int LB_M = 10;        // history length (scipy default `maxcor`)
constexpr int LB_2M = 2 * LB_M;
constexpr int LB_P = LB_2M + 1;  // odd pitch of the dense 2m x 2m matrices: lane = row accesses are bank-conflict free

enum { LB_FG = 0, LB_NEW_X = 1, LB_CONVERGED = 2, LB_STOPPED = 3, LB_ABNORMAL = 4 };
enum { LB_PH_START = 0, LB_PH_LINESEARCH = 1 };
// messages (status words of the reference's OptimizeResult)
enum { LB_MSG_NONE = 0, LB_MSG_PGTOL = 401, LB_MSG_FACTR = 402, LB_MSG_MAXFUN = 502, LB_MSG_MAXITER = 504, LB_MSG_ABNORMAL = 800 };

// double scalars of one problem
enum { S_F = 0, S_FOLD, S_THETA, S_DTD, S_GDOLD, S_STP, S_STPMX, S_STX, S_FX, S_GX, S_STY, S_FY, S_GY, S_STMIN, S_STMAX,
       S_WIDTH, S_WIDTH1, S_FINIT, S_GINIT, S_GTEST, S_NSCALARS = 24 };
// int scalars of one problem
enum { I_TASK = 0, I_PHASE, I_COL, I_HEAD, I_ITER, I_NFEV, I_BRACKT, I_STAGE, I_IFUN, I_IBACK, I_MSG, I_NINTS = 16 };

struct LbLayout {
  int64_t N;
  int D;
  // per-problem strides (in doubles) inside the state buffer
  size_t off_scal, off_ss, off_sy, off_g, off_d, off_z, off_t, off_gold, off_wk1, off_wk2, off_S, off_Y, per_problem;
  size_t ints_offset_bytes;   // where the int block starts inside the state buffer
};

__host__ __device__ inline LbLayout lb_layout(int64_t N, int D) {
  LbLayout L;
  L.N = N; L.D = D;
  size_t o = 0;
  L.off_scal = o; o += S_NSCALARS;
  L.off_ss = o; o += LB_M * LB_M;
  L.off_sy = o; o += LB_M * LB_M;
  const size_t Dp = (size_t)((D + 1) & ~1);
  L.off_g = o; o += Dp;
  L.off_d = o; o += Dp;
  L.off_z = o; o += Dp;
  L.off_t = o; o += Dp;
  L.off_gold = o; o += Dp;
  L.off_wk1 = o; o += Dp;
  L.off_wk2 = o; o += Dp;
  L.off_S = o; o += (size_t)LB_M * Dp;
  L.off_Y = o; o += (size_t)LB_M * Dp;
  L.per_problem = o;
  L.ints_offset_bytes = (size_t)N * L.per_problem * sizeof(double);
  return L;
}

// ---- CTA-wide reductions (fixed order: lane tree, then warps 0..3) ------------------------------------------------------
__device__ __forceinline__ double lb_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double lb_block_sum(double v, double* red) {
  v = lb_warp_sum(v);
  __syncthreads();   // red may still be read from a previous reduction
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < LB_WARPS; w++) s += red[w];
  return s;
}
__device__ __forceinline__ double lb_block_max(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = red[0];
#pragma unroll
  for (int w = 1; w < LB_WARPS; w++) s = fmax(s, red[w]);
  return s;
}
// (min value, lowest index attaining it) over the CTA
__device__ __forceinline__ void lb_block_argmin(double v, int idx, double* red, int* redi, double& vout, int& iout) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double v2 = __shfl_xor_sync(0xffffffffu, v, o);
    const int i2 = __shfl_xor_sync(0xffffffffu, idx, o);
    if (v2 < v || (v2 == v && i2 < idx)) { v = v2; idx = i2; }
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = v; redi[threadIdx.x >> 5] = idx; }
  __syncthreads();
  vout = red[0]; iout = redi[0];
#pragma unroll
  for (int w = 1; w < LB_WARPS; w++)
    if (red[w] < vout || (red[w] == vout && redi[w] < iout)) { vout = red[w]; iout = redi[w]; }
}

// K dot products at once: out[k] = sum_i term(k, i); warps take k = warp, warp + 4, ...; lanes stride over i.
template <class F>
__device__ __forceinline__ void lb_multi_dot(int K, int D, double* out, F term) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = warp; k < K; k += LB_WARPS) {
    double acc = 0.0;
    for (int i = lane; i < D; i += 32) acc += term(k, i);
    acc = lb_warp_sum(acc);
    if (lane == 0) out[k] = acc;
  }
  __syncthreads();
}

// ---- dense (<= 20 x 20) LU with partial pivoting, warp 0 (all 32 lanes), operands in shared memory ----------------------
// Factorises A (row-major, pitch LB_2M) in place, LAPACK style (whole rows are interchanged); piv[i] = row swapped into
// position i.  Lane = row during the elimination, lane = column during the interchange.  Returns false on a zero pivot.
__device__ bool lb_lu_factor_w(double* A, int* piv, int k) {
  const int lane = threadIdx.x & 31;
  for (int i = 0; i < k; i++) {
    double a = (lane >= i && lane < k) ? fabs(A[lane * LB_P + i]) : -1.0;
    int idx = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {   // largest magnitude, lowest row on ties (first maximum)
      const double a2 = __shfl_xor_sync(0xffffffffu, a, o);
      const int i2 = __shfl_xor_sync(0xffffffffu, idx, o);
      if (a2 > a || (a2 == a && i2 < idx)) { a = a2; idx = i2; }
    }
    if (!(a > 0.0)) return false;
    if (lane == 0) piv[i] = idx;
    if (idx != i && lane < k) { const double tmp = A[i * LB_P + lane]; A[i * LB_P + lane] = A[idx * LB_P + lane]; A[idx * LB_P + lane] = tmp; }
    __syncwarp();
    if (lane > i && lane < k) {
      const double fct = A[lane * LB_P + i] / A[i * LB_P + i];
      A[lane * LB_P + i] = fct;
      for (int c = i + 1; c < k; c++) A[lane * LB_P + c] -= fct * A[i * LB_P + c];
    }
    __syncwarp();
  }
  return true;
}
// x = A^{-1} b from the factors (b, x, ys: shared memory; ys is scratch of LB_2M doubles).  All interchanges are applied to
// the right-hand side first, then the unit-lower and the upper solve run column by column (lane = row).
__device__ void lb_lu_solve_w(const double* LU, const int* piv, int k, const double* b, double* x, double* ys) {
  const int lane = threadIdx.x & 31;
  if (lane < k) ys[lane] = b[lane];
  __syncwarp();
  if (lane == 0)
    for (int i = 0; i < k; i++) { const int p = piv[i]; if (p != i) { const double tmp = ys[i]; ys[i] = ys[p]; ys[p] = tmp; } }
  __syncwarp();
  for (int i = 0; i < k; i++) {
    const double yi = ys[i];
    if (lane > i && lane < k) ys[lane] -= LU[lane * LB_P + i] * yi;
    __syncwarp();
  }
  for (int c = k - 1; c >= 0; c--) {
    if (lane == c) x[c] = ys[c] / LU[c * LB_P + c];
    __syncwarp();
    const double xc = x[c];
    if (lane < c) ys[lane] -= LU[lane * LB_P + c] * xc;
    __syncwarp();
  }
}

// ---- More'-Thuente line search (MINPACK-2 dcstep / dcsrch), thread 0, state in the problem's scalar block ---------------
struct LbLs { double stx, fx, gx, sty, fy, gy, stp, stmin, stmax, width, width1, finit, ginit, gtest, stpmx; int brackt, stage; };

__device__ void lb_dcstep(LbLs& ls, double fp, double dp) {
  double stx = ls.stx, fx = ls.fx, dx = ls.gx, sty = ls.sty, fy = ls.fy, dy = ls.gy, stp = ls.stp;
  int brackt = ls.brackt;
  const double stpmin = ls.stmin, stpmax = ls.stmax;
  const double sgnd = dp * (dx / fabs(dx));
  double stpf;
  if (fp > fx) {
    const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    const double s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
    double gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
    if (stp < stx) gamma = -gamma;
    const double p = (gamma - dx) + theta, q = ((gamma - dx) + gamma) + dp, r = p / q;
    const double stpc = stx + r * (stp - stx);
    const double stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
    stpf = (fabs(stpc - stx) < fabs(stpq - stx)) ? stpc : stpc + (stpq - stpc) / 2.0;
    brackt = 1;
  } else if (sgnd < 0.0) {
    const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    const double s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
    double gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
    if (stp > stx) gamma = -gamma;
    const double p = (gamma - dp) + theta, q = ((gamma - dp) + gamma) + dx, r = p / q;
    const double stpc = stp + r * (stx - stp);
    const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
    stpf = (fabs(stpc - stp) > fabs(stpq - stp)) ? stpc : stpq;
    brackt = 1;
  } else if (fabs(dp) < fabs(dx)) {
    const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    const double s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
    double gamma = s * sqrt(fmax(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
    if (stp > stx) gamma = -gamma;
    const double p = (gamma - dp) + theta, q = (gamma + (dx - dp)) + gamma, r = p / q;
    double stpc;
    if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
    else if (stp > stx) stpc = stpmax;
    else stpc = stpmin;
    const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
    if (brackt) {
      stpf = (fabs(stpc - stp) < fabs(stpq - stp)) ? stpc : stpq;
      if (stp > stx) stpf = fmin(stp + 0.66 * (sty - stp), stpf);
      else stpf = fmax(stp + 0.66 * (sty - stp), stpf);
    } else {
      stpf = (fabs(stpc - stp) > fabs(stpq - stp)) ? stpc : stpq;
      stpf = fmin(stpmax, stpf);
      stpf = fmax(stpmin, stpf);
    }
  } else {
    if (brackt) {
      const double theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
      const double s = fmax(fabs(theta), fmax(fabs(dy), fabs(dp)));
      double gamma = s * sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
      if (stp > sty) gamma = -gamma;
      const double p = (gamma - dp) + theta, q = ((gamma - dp) + gamma) + dy, r = p / q;
      stpf = stp + r * (sty - stp);
    } else if (stp > stx) stpf = stpmax;
    else stpf = stpmin;
  }
  if (fp > fx) { sty = stp; fy = fp; dy = dp; }
  else {
    if (sgnd < 0.0) { sty = stx; fy = fx; dy = dx; }
    stx = stp; fx = fp; dx = dp;
  }
  ls.stx = stx; ls.fx = fx; ls.gx = dx; ls.sty = sty; ls.fy = fy; ls.gy = dy; ls.stp = stpf; ls.brackt = brackt;
}

enum { LS_FG = 0, LS_CONVERGENCE = 1, LS_WARNING = 2, LS_ERROR = 3 };
constexpr double LB_FTOL = 1.0e-3, LB_GTOL = 0.9, LB_XTOL = 0.1, LB_BIG = 1.0e10;

__device__ int lb_dcsrch(LbLs& ls, double f, double g, bool start) {
  const double stpmin = 0.0, stpmax = ls.stpmx;
  if (start) {
    if (ls.stp < stpmin || ls.stp > stpmax || g >= 0.0) return LS_ERROR;
    ls.brackt = 0; ls.stage = 1; ls.finit = f; ls.ginit = g; ls.gtest = LB_FTOL * g;
    ls.width = stpmax - stpmin; ls.width1 = 2.0 * (stpmax - stpmin);
    ls.stx = 0.0; ls.fx = f; ls.gx = g; ls.sty = 0.0; ls.fy = f; ls.gy = g; ls.stmin = 0.0; ls.stmax = ls.stp + 4.0 * ls.stp;
    return LS_FG;
  }
  const double stp = ls.stp;
  const double ftest = ls.finit + stp * ls.gtest;
  if (ls.stage == 1 && f <= ftest && g >= 0.0) ls.stage = 2;
  if (ls.brackt && (stp <= ls.stmin || stp >= ls.stmax)) return LS_WARNING;
  if (ls.brackt && ls.stmax - ls.stmin <= LB_XTOL * ls.stmax) return LS_WARNING;
  if (stp == stpmax && f <= ftest && g <= ls.gtest) return LS_WARNING;
  if (stp == stpmin && (f > ftest || g >= ls.gtest)) return LS_WARNING;
  if (f <= ftest && fabs(g) <= LB_GTOL * (-ls.ginit)) return LS_CONVERGENCE;
  if (ls.stage == 1 && f <= ls.fx && f > ftest) {
    const double gt = ls.gtest;
    const double fm = f - stp * gt, gm = g - gt;
    ls.fx -= ls.stx * gt; ls.fy -= ls.sty * gt; ls.gx -= gt; ls.gy -= gt;
    lb_dcstep(ls, fm, gm);
    ls.fx += ls.stx * gt; ls.fy += ls.sty * gt; ls.gx += gt; ls.gy += gt;
  } else {
    lb_dcstep(ls, f, g);
  }
  if (ls.brackt) {
    if (fabs(ls.sty - ls.stx) >= 0.66 * ls.width1) ls.stp = ls.stx + 0.5 * (ls.sty - ls.stx);
    ls.width1 = ls.width;
    ls.width = fabs(ls.sty - ls.stx);
  }
  if (ls.brackt) { ls.stmin = fmin(ls.stx, ls.sty); ls.stmax = fmax(ls.stx, ls.sty); }
  else { ls.stmin = ls.stp + 1.1 * (ls.stp - ls.stx); ls.stmax = ls.stp + 4.0 * (ls.stp - ls.stx); }
  ls.stp = fmin(fmax(ls.stp, stpmin), stpmax);
  if ((ls.brackt && (ls.stp <= ls.stmin || ls.stp >= ls.stmax)) || (ls.brackt && ls.stmax - ls.stmin <= LB_XTOL * ls.stmax))
    ls.stp = ls.stx;
  return LS_FG;
}

__device__ __forceinline__ void lb_load_ls(LbLs& ls, const double* sc, const int* is) {
  ls.stx = sc[S_STX]; ls.fx = sc[S_FX]; ls.gx = sc[S_GX]; ls.sty = sc[S_STY]; ls.fy = sc[S_FY]; ls.gy = sc[S_GY];
  ls.stp = sc[S_STP]; ls.stmin = sc[S_STMIN]; ls.stmax = sc[S_STMAX]; ls.width = sc[S_WIDTH]; ls.width1 = sc[S_WIDTH1];
  ls.finit = sc[S_FINIT]; ls.ginit = sc[S_GINIT]; ls.gtest = sc[S_GTEST]; ls.stpmx = sc[S_STPMX];
  ls.brackt = is[I_BRACKT]; ls.stage = is[I_STAGE];
}
__device__ __forceinline__ void lb_store_ls(const LbLs& ls, double* sc, int* is) {
  sc[S_STX] = ls.stx; sc[S_FX] = ls.fx; sc[S_GX] = ls.gx; sc[S_STY] = ls.sty; sc[S_FY] = ls.fy; sc[S_GY] = ls.gy;
  sc[S_STP] = ls.stp; sc[S_STMIN] = ls.stmin; sc[S_STMAX] = ls.stmax; sc[S_WIDTH] = ls.width; sc[S_WIDTH1] = ls.width1;
  sc[S_FINIT] = ls.finit; sc[S_GINIT] = ls.ginit; sc[S_GTEST] = ls.gtest; sc[S_STPMX] = ls.stpmx;
  is[I_BRACKT] = ls.brackt; is[I_STAGE] = ls.stage;
}

struct LbParams {
  double factr, pgtol, sign;
  int maxiter, maxfun, maxls;
};

// shared-memory working set of one CTA
struct LbShared {
  double red[LB_WARPS];
  int redi[LB_WARPS];
  double sc[S_NSCALARS];
  int is[I_NINTS];
  double SS[LB_M * LB_M], SY[LB_M * LB_M];
  double Minv[LB_2M * LB_P];   // LU of M^{-1}
  double K3[LB_2M * LB_P];
  int pivM[LB_2M], pivK[LB_2M];
  double p[LB_2M], c[LB_2M], v[LB_2M], wbp[LB_2M], mc[LB_2M], rhs[LB_2M], ys[LB_2M];
  double dots[LB_2M * (LB_2M + 1) / 2 + LB_2M];
  double bc[8];   // broadcast scalars
  int bi[8];
};

// logical history row j (0 = oldest) -> physical row of the ring
__device__ __forceinline__ int lb_row(int head, int j) { return (head + j) % LB_M; }

// projected-gradient infinity norm
__device__ double lb_projgr(int D, const double* x, const double* g, const double* l, const double* u, double* red) {
  double mx = 0.0;
  for (int i = threadIdx.x; i < D; i += LB_THREADS) {
    const double gi = g[i];
    const double pg = (gi < 0.0) ? fmax(x[i] - u[i], gi) : fmin(x[i] - l[i], gi);
    mx = fmax(mx, fabs(pg));
  }
  return lb_block_max(mx, red);
}

// M^{-1}[a][b] of the compact representation from S'S, S'Y and theta
__device__ __forceinline__ double lb_minv_entry(const LbShared& sh, int c, double theta, int a, int b) {
  if (a < c && b < c) return (a == b) ? -sh.SY[a * LB_M + a] : 0.0;
  if (a < c) return ((b - c) > a) ? sh.SY[(b - c) * LB_M + a] : 0.0;          // L^T[a][b-c] = L[b-c][a], b-c > a
  if (b < c) return ((a - c) > b) ? sh.SY[(a - c) * LB_M + b] : 0.0;          // L[a-c][b], strictly lower
  return theta * sh.SS[(a - c) * LB_M + (b - c)];
}

// builds LU(M^{-1}) in sh.Minv; returns false when singular.  All threads call.
__device__ bool lb_factor_minv(LbShared& sh, int c, double theta) {
  __syncthreads();
  const int k = 2 * c;
  for (int e = threadIdx.x; e < k * k; e += LB_THREADS) {
    const int a = e / k, b = e - a * k;
    sh.Minv[a * LB_P + b] = lb_minv_entry(sh, c, theta, a, b);
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const bool ok = lb_lu_factor_w(sh.Minv, sh.pivM, k);
    if (threadIdx.x == 0) sh.bi[0] = ok ? 1 : 0;
  }
  __syncthreads();
  return sh.bi[0] != 0;
}

// W[j][i] of the compact representation: j < c -> Y_j[i], else theta * S_{j-c}[i].  `Wsm` != nullptr: the 2c x D matrix has
// been staged in shared memory for this iteration (every element is read 2c + 3 times by the Cauchy / subspace steps, and
// the 2c (2c + 1) / 2 masked Gram dot products are latency-bound when each operand is an L2 access).
__device__ __forceinline__ double lb_W(const double* S, const double* Y, size_t Dp, int head, int c, double theta, int j, int i,
                                       const double* Wsm = nullptr) {
  if (Wsm != nullptr) return Wsm[(size_t)j * Dp + i];
  return (j < c) ? Y[(size_t)lb_row(head, j) * Dp + i] : theta * S[(size_t)lb_row(head, j - c) * Dp + i];
}

__global__ void __launch_bounds__(LB_THREADS)
lbfgsb_step_kernel(LbLayout L, double* __restrict__ X, const double* __restrict__ f_in, const double* __restrict__ g_in,
                   const double* __restrict__ lower, const double* __restrict__ upper, double* __restrict__ state, LbParams prm,
                   int32_t* __restrict__ n_active, int stage_w) {
  __shared__ LbShared sh;
  extern __shared__ __align__(16) double lb_dyn[];   // [2 m][Dp] staged W (when it fits: stage_w != 0)
  const int64_t b = blockIdx.x;
  const int D = L.D, tid = threadIdx.x;
  const size_t Dp = (size_t)((D + 1) & ~1);
  double* ps = state + (size_t)b * L.per_problem;
  int* pi = reinterpret_cast<int*>(reinterpret_cast<char*>(state) + L.ints_offset_bytes) + (size_t)b * I_NINTS;
  if (pi[I_TASK] != LB_FG) return;   // finished problems keep their X row
  double* x = X + (size_t)b * D;
  double* g = ps + L.off_g; double* d = ps + L.off_d; double* z = ps + L.off_z; double* t = ps + L.off_t;
  double* gold = ps + L.off_gold; double* wk1 = ps + L.off_wk1; double* wk2 = ps + L.off_wk2;
  double* S = ps + L.off_S; double* Y = ps + L.off_Y;
  const double* l = lower; const double* u = upper;

  if (tid < S_NSCALARS) sh.sc[tid] = ps[L.off_scal + tid];
  if (tid < I_NINTS) sh.is[tid] = pi[tid];
  for (int i = tid; i < LB_M * LB_M; i += LB_THREADS) { sh.SS[i] = ps[L.off_ss + i]; sh.SY[i] = ps[L.off_sy + i]; }
  // the new function value and gradient (the optimiser minimises sign * f_in)
  for (int i = tid; i < D; i += LB_THREADS) g[i] = prm.sign * g_in[(size_t)b * D + i];
  __syncthreads();
  if (tid == 0) { sh.sc[S_F] = prm.sign * f_in[b]; sh.is[I_NFEV] += 1; }
  __syncthreads();

  bool begin_iteration = false, ls_failed = false;
  if (sh.is[I_PHASE] == LB_PH_START) {
    const double sbg = lb_projgr(D, x, g, l, u, sh.red);
    if (sbg <= prm.pgtol) {
      if (tid == 0) { sh.is[I_TASK] = LB_CONVERGED; sh.is[I_MSG] = LB_MSG_PGTOL; }
    } else begin_iteration = true;
  } else {
    // ---- line search continues: judge the trial point
    double part = 0.0;
    for (int i = tid; i < D; i += LB_THREADS) part += g[i] * d[i];
    const double gd = lb_block_sum(part, sh.red);
    if (tid == 0) {
      LbLs ls; lb_load_ls(ls, sh.sc, sh.is);
      const int res = lb_dcsrch(ls, sh.sc[S_F], gd, false);
      lb_store_ls(ls, sh.sc, sh.is);
      int action;   // 0 take another trial, 1 failed, 2 new iterate
      if (res == LS_FG) action = (sh.is[I_IBACK] + 1 >= prm.maxls) ? 1 : 0;
      else if (res == LS_ERROR) action = 1;
      else action = 2;
      sh.bi[1] = action;
    }
    __syncthreads();
    const int action = sh.bi[1];
    if (action == 0) {
      // ---- next trial point of the same line search
      if (tid == 0) { sh.is[I_IFUN] += 1; sh.is[I_IBACK] = sh.is[I_IFUN] - 1; }
      const double stp = sh.sc[S_STP];
      for (int i = tid; i < D; i += LB_THREADS) x[i] = (stp == 1.0) ? z[i] : stp * d[i] + t[i];
    } else if (action == 1) {
      ls_failed = true;
    } else {
      // ---- a new iterate: stopping tests, then the limited-memory update
      const double stp = sh.sc[S_STP];
      const double sbg = lb_projgr(D, x, g, l, u, sh.red);
      if (tid == 0) {
        sh.is[I_ITER] += 1;
        const double f = sh.sc[S_F], fold = sh.sc[S_FOLD];
        int task = LB_FG, msg = LB_MSG_NONE;
        if (sh.is[I_ITER] >= prm.maxiter) { task = LB_STOPPED; msg = LB_MSG_MAXITER; }
        else if (sh.is[I_NFEV] > prm.maxfun) { task = LB_STOPPED; msg = LB_MSG_MAXFUN; }
        else if (sbg <= prm.pgtol) { task = LB_CONVERGED; msg = LB_MSG_PGTOL; }
        else {
          const double ddum = fmax(fabs(fold), fmax(fabs(f), 1.0));
          if ((fold - f) <= DBL_EPSILON * prm.factr * ddum) { task = LB_CONVERGED; msg = LB_MSG_FACTR; }
        }
        sh.is[I_TASK] = task; sh.is[I_MSG] = msg;
      }
      __syncthreads();
      if (sh.is[I_TASK] == LB_FG) {
        // y = g - gold (into wk1), rr = y.y
        double prr = 0.0;
        for (int i = tid; i < D; i += LB_THREADS) { const double yi = g[i] - gold[i]; wk1[i] = yi; prr += yi * yi; }
        const double rr = lb_block_sum(prr, sh.red);
        const double gdold = sh.sc[S_GDOLD];
        double dr, ddum;
        if (stp == 1.0) { dr = gd - gdold; ddum = -gdold; }
        else { dr = (gd - gdold) * stp; ddum = -gdold * stp; }
        if (dr > DBL_EPSILON * ddum) {
          int c = sh.is[I_COL], head = sh.is[I_HEAD];
          if (c == LB_M) {  // drop the oldest pair: advance the ring, shift the small matrices
            __syncthreads();
            if (tid == 0) {
              for (int a = 0; a < LB_M - 1; a++)
                for (int bb = 0; bb < LB_M - 1; bb++) {
                  sh.SS[a * LB_M + bb] = sh.SS[(a + 1) * LB_M + bb + 1];
                  sh.SY[a * LB_M + bb] = sh.SY[(a + 1) * LB_M + bb + 1];
                }
            }
            head = (head + 1) % LB_M;
            c -= 1;
            __syncthreads();
          }
          const int prow = lb_row(head, c);
          for (int i = tid; i < D; i += LB_THREADS) {
            S[(size_t)prow * Dp + i] = (stp == 1.0) ? d[i] : stp * d[i];
            Y[(size_t)prow * Dp + i] = wk1[i];
          }
          __syncthreads();
          // new row / column of S'S and S'Y: 3 (c + 1) dot products
          const double* snew = S + (size_t)prow * Dp;
          const double* ynew = Y + (size_t)prow * Dp;
          lb_multi_dot(3 * (c + 1), D, sh.dots, [&](int k, int i) {
            const int j = k / 3, which = k - 3 * j;
            const size_t r = (size_t)lb_row(head, j) * Dp;
            if (which == 0) return S[r + i] * snew[i];        // SS[c][j]
            if (which == 1) return snew[i] * Y[r + i];        // SY[c][j]
            return S[r + i] * ynew[i];                        // SY[j][c]
          });
          if (tid == 0) {
            for (int j = 0; j <= c; j++) {
              sh.SS[c * LB_M + j] = sh.SS[j * LB_M + c] = sh.dots[3 * j];
              sh.SY[c * LB_M + j] = sh.dots[3 * j + 1];
              sh.SY[j * LB_M + c] = sh.dots[3 * j + 2];
            }
            const double dtd = sh.sc[S_DTD];
            sh.SS[c * LB_M + c] = (stp == 1.0) ? dtd : stp * stp * dtd;
            sh.SY[c * LB_M + c] = dr;
            sh.is[I_COL] = c + 1; sh.is[I_HEAD] = head;
            sh.sc[S_THETA] = rr / dr;
          }
          __syncthreads();
        }
        begin_iteration = true;
      }
    }
  }

  // ---- set up iterations until a trial point is out (at most twice: a failed line search restarts without memory)
  for (int attempt = 0; attempt < 3 && (begin_iteration || ls_failed); attempt++) {
    if (ls_failed) {
      ls_failed = false;
      for (int i = tid; i < D; i += LB_THREADS) { x[i] = t[i]; g[i] = gold[i]; }
      __syncthreads();
      if (tid == 0) {
        sh.sc[S_F] = sh.sc[S_FOLD];
        if (sh.is[I_COL] == 0) { sh.is[I_TASK] = LB_ABNORMAL; sh.is[I_MSG] = LB_MSG_ABNORMAL; }
        else { sh.is[I_COL] = 0; sh.is[I_HEAD] = 0; sh.sc[S_THETA] = 1.0; }
      }
      __syncthreads();
      if (sh.is[I_TASK] == LB_ABNORMAL) { begin_iteration = false; break; }
      begin_iteration = true;
    }
    begin_iteration = false;
    // ================= generalised Cauchy point (oracle/lbfgsb.py::_cauchy) =================
    bool ok = true;
    double* dC = t;        // Cauchy direction (t and gold are re-saved at the end of the set-up)
    double* tbreak = gold;
    double* freem = wk2;   // 1.0 = free at the Cauchy point
    double* xcp = z;
    for (int retry = 0; retry < 2; retry++) {
      ok = true;
      const int c = sh.is[I_COL], head = sh.is[I_HEAD];
      const double theta = sh.sc[S_THETA];
      const double* Wsm = nullptr;
      if (stage_w && c > 0) {
        for (int e = tid; e < 2 * c * D; e += LB_THREADS) {
          const int j = e / D, i = e - j * D;
          lb_dyn[(size_t)j * Dp + i] = lb_W(S, Y, Dp, head, c, theta, j, i);
        }
        Wsm = lb_dyn;
        __syncthreads();
      }
      int nb = 0;
      double pf1 = 0.0;
      for (int i = tid; i < D; i += LB_THREADS) {
        const double xi = x[i], tl = xi - l[i], tu = u[i] - xi, neg = -g[i];
        const bool fixed = ((tl <= 0.0) && (neg <= 0.0)) || ((tu <= 0.0) && (neg >= 0.0) && !(tl <= 0.0)) || (neg == 0.0);
        const double di = fixed ? 0.0 : neg;
        double tb = CUDART_INF_F;
        if (!fixed && neg < 0.0 && isfinite(l[i])) tb = tl / (-neg);
        else if (!fixed && neg > 0.0 && isfinite(u[i])) tb = tu / neg;
        dC[i] = di; tbreak[i] = tb; freem[i] = fixed ? 0.0 : 1.0; xcp[i] = xi;
        nb += isfinite(tb) ? 1 : 0;
        pf1 += di * di;
      }
      const double f1_0 = -lb_block_sum(pf1, sh.red);
      const int nbreak = (int)(lb_block_sum((double)nb, sh.red) + 0.5);
      // p = W dC
      if (c > 0) {
        lb_multi_dot(2 * c, D, sh.p, [&](int k, int i) { return lb_W(S, Y, Dp, head, c, theta, k, i, Wsm) * dC[i]; });
        if (!lb_factor_minv(sh, c, theta)) { ok = false; }
      }
      if (ok) {
        if (tid < 32 && c > 0) lb_lu_solve_w(sh.Minv, sh.pivM, 2 * c, sh.p, sh.v, sh.ys);
        if (tid == 0) {
          double f1 = f1_0, f2 = -theta * f1;
          for (int k = 0; k < 2 * c; k++) sh.c[k] = 0.0;
          if (c > 0) {
            double pv = 0.0;
            for (int k = 0; k < 2 * c; k++) pv += sh.p[k] * sh.v[k];
            f2 -= pv;
          }
          sh.bc[0] = f1; sh.bc[1] = f2; sh.bc[2] = f2;      // f1, f2, f2_org
          sh.bc[3] = -f1 / f2;                              // dtm
          sh.bc[4] = 0.0; sh.bc[5] = 0.0;                   // tsum, tj0
          sh.bi[2] = 0;                                     // all_fixed_exit
        }
        __syncthreads();
        for (int kbp = 0; kbp < nbreak; kbp++) {
          // next breakpoint: smallest remaining t (lowest index on ties = stable sort order)
          double vmin = CUDART_INF_F; int imin = 0x7fffffff;
          for (int i = tid; i < D; i += LB_THREADS) { const double tb = tbreak[i]; if (tb < vmin) { vmin = tb; imin = i; } }
          double tj; int bidx;
          lb_block_argmin(vmin, imin, sh.red, sh.redi, tj, bidx);
          const double dt = tj - sh.bc[5];
          if (sh.bc[3] < dt) break;
          // consume the breakpoint
          if (c > 0 && tid < 2 * c) sh.wbp[tid] = lb_W(S, Y, Dp, head, c, theta, tid, bidx, Wsm);
          __syncthreads();
          if (tid < 32) {
            // lane 0 moves the variable to its bound; the 2c-sized algebra (c += dt p, v = M wbp) runs on the warp
            if (tid == 0) {
              sh.bc[4] += dt;
              const double dibp = dC[bidx];
              dC[bidx] = 0.0;
              double zibp;
              if (dibp > 0.0) { zibp = u[bidx] - x[bidx]; xcp[bidx] = u[bidx]; }
              else { zibp = l[bidx] - x[bidx]; xcp[bidx] = l[bidx]; }
              freem[bidx] = 0.0;
              tbreak[bidx] = CUDART_INF_F;
              sh.bc[6] = dibp; sh.bc[7] = zibp;
              sh.bi[2] = (kbp == nbreak - 1 && nbreak == D) ? 1 : 0;
            }
            __syncwarp();
            const bool last = sh.bi[2] != 0;
            if (!last && c > 0) {
              if (tid < 2 * c) sh.c[tid] += dt * sh.p[tid];
              __syncwarp();
              lb_lu_solve_w(sh.Minv, sh.pivM, 2 * c, sh.wbp, sh.v, sh.ys);
            }
            if (tid == 0) {
              double f1 = sh.bc[0], f2 = sh.bc[1], dtm = sh.bc[3];
              const double dibp = sh.bc[6], zibp = sh.bc[7];
              if (last) {
                dtm = dt;
              } else {
                const double dibp2 = dibp * dibp;
                f1 = f1 + dt * f2 + dibp2 - theta * dibp * zibp;
                f2 = f2 - theta * dibp2;
                if (c > 0) {
                  double wmc = 0.0, wmp = 0.0, wmw = 0.0;
                  for (int k = 0; k < 2 * c; k++) { wmc += sh.c[k] * sh.v[k]; wmp += sh.p[k] * sh.v[k]; wmw += sh.wbp[k] * sh.v[k]; }
                  for (int k = 0; k < 2 * c; k++) sh.p[k] -= dibp * sh.wbp[k];
                  f1 += dibp * wmc;
                  f2 += 2.0 * dibp * wmp - dibp2 * wmw;
                }
                f2 = fmax(DBL_EPSILON * sh.bc[2], f2);
                dtm = -f1 / f2;
                sh.bc[5] = tj;
              }
              sh.bc[0] = f1; sh.bc[1] = f2; sh.bc[3] = dtm;
            }
          }
          __syncthreads();
          if (sh.bi[2]) break;
        }
        __syncthreads();
        if (tid == 0 && !sh.bi[2]) { sh.bc[3] = fmax(sh.bc[3], 0.0); sh.bc[4] += sh.bc[3]; }
        __syncthreads();
        if (!sh.bi[2]) {
          const double tsum = sh.bc[4];
          for (int i = tid; i < D; i += LB_THREADS) if (dC[i] != 0.0) xcp[i] = x[i] + tsum * dC[i];
        }
        if (tid == 0 && c > 0) for (int k = 0; k < 2 * c; k++) sh.c[k] += sh.bc[3] * sh.p[k];
        __syncthreads();
        // ================= subspace minimisation (oracle/lbfgsb.py::_subsm) =================
        double pnf = 0.0;
        for (int i = tid; i < D; i += LB_THREADS) pnf += freem[i];
        const int nfree = (int)(lb_block_sum(pnf, sh.red) + 0.5);
        if (nfree > 0 && c > 0) {
          if (tid < 32) lb_lu_solve_w(sh.Minv, sh.pivM, 2 * c, sh.c, sh.mc, sh.ys);
          __syncthreads();
          // r (into wk1, zero on fixed variables)
          for (int i = tid; i < D; i += LB_THREADS) {
            double r = 0.0;
            if (freem[i] != 0.0) {
              double wm = 0.0;
              for (int k = 0; k < 2 * c; k++) wm += lb_W(S, Y, Dp, head, c, theta, k, i, Wsm) * sh.mc[k];
              r = -theta * (xcp[i] - x[i]) - g[i] + wm;
            }
            wk1[i] = r;
          }
          __syncthreads();
          // Wz Wz^T (upper triangle) and Wz r
          const int k2 = 2 * c, npair = k2 * (k2 + 1) / 2;
          lb_multi_dot(npair + k2, D, sh.dots, [&](int k, int i) {
            if (freem[i] == 0.0) return 0.0;
            if (k >= npair) return lb_W(S, Y, Dp, head, c, theta, k - npair, i, Wsm) * wk1[i];
            int a = 0, rem = k;   // k -> (a, bcol) with a <= bcol, rows of length k2 - a
            while (rem >= k2 - a) { rem -= k2 - a; a++; }
            const int bcol = a + rem;
            return lb_W(S, Y, Dp, head, c, theta, a, i, Wsm) * lb_W(S, Y, Dp, head, c, theta, bcol, i, Wsm);
          });
          // K3 = M^{-1} - Wz Wz^T / theta (sh.Minv holds the LU of M^{-1}, so its entries are rebuilt)
          for (int e = tid; e < k2 * k2; e += LB_THREADS) {
            const int a = e / k2, bcol = e - a * k2;
            const int lo = a < bcol ? a : bcol, hi = a < bcol ? bcol : a;
            const int kk = lo * k2 - lo * (lo - 1) / 2 + (hi - lo);   // index of (lo, hi) in the packed upper triangle
            sh.K3[a * LB_P + bcol] = lb_minv_entry(sh, c, theta, a, bcol) - sh.dots[kk] / theta;
          }
          if (tid < k2) sh.rhs[tid] = sh.dots[npair + tid];
          __syncthreads();
          if (tid < 32) {
            const bool okf = lb_lu_factor_w(sh.K3, sh.pivK, k2);
            if (okf) lb_lu_solve_w(sh.K3, sh.pivK, k2, sh.rhs, sh.v, sh.ys);
            if (tid == 0) sh.bi[3] = okf ? 1 : 0;
          }
          __syncthreads();
          if (!sh.bi[3]) ok = false;
          if (ok) {
            // Newton step on the free variables (into wk1), projected point (into d as scratch), bound hits
            int hit = 0;
            double pdd = 0.0;
            for (int i = tid; i < D; i += LB_THREADS) {
              double xb = xcp[i];
              if (freem[i] != 0.0) {
                double wv = 0.0;
                for (int k = 0; k < k2; k++) wv += lb_W(S, Y, Dp, head, c, theta, k, i, Wsm) * sh.v[k];
                const double dF = (wk1[i] + wv / theta) / theta;
                wk1[i] = dF;
                xb = fmin(fmax(xcp[i] + dF, l[i]), u[i]);
                if (xb == l[i] || xb == u[i]) hit = 1;
              }
              d[i] = xb;
              pdd += (xb - x[i]) * g[i];
            }
            const double ddp = lb_block_sum(pdd, sh.red);
            const int any_hit = (int)(lb_block_sum((double)hit, sh.red) > 0.5);
            if (any_hit && ddp > 0.0) {
              // truncate the Newton step at the first bound (lowest index on ties)
              double amin = 1.0; int imin = 0x7fffffff;
              for (int i = tid; i < D; i += LB_THREADS) {
                if (freem[i] == 0.0) continue;
                const double dk = wk1[i];
                double t1 = 2.0;
                if (dk < 0.0 && isfinite(l[i])) { const double t2 = l[i] - xcp[i]; t1 = (t2 >= 0.0) ? 0.0 : t2 / dk; }
                else if (dk > 0.0 && isfinite(u[i])) { const double t2 = u[i] - xcp[i]; t1 = (t2 <= 0.0) ? 0.0 : t2 / dk; }
                if (t1 < amin || (t1 == amin && t1 < 1.0 && i < imin)) { amin = t1; imin = i; }
              }
              double alpha; int ibd;
              lb_block_argmin(amin, imin, sh.red, sh.redi, alpha, ibd);
              for (int i = tid; i < D; i += LB_THREADS) {
                double xb = xcp[i];
                if (freem[i] != 0.0) {
                  xb = xcp[i] + alpha * wk1[i];
                  if (alpha < 1.0 && i == ibd) xb = (wk1[i] > 0.0) ? u[i] : l[i];
                }
                z[i] = xb;
              }
            } else {
              for (int i = tid; i < D; i += LB_THREADS) z[i] = d[i];
            }
            __syncthreads();
          }
        }
      }
      if (ok) break;
      // singular compact matrices: forget the history and redo the set-up
      __syncthreads();
      if (tid == 0) { sh.is[I_COL] = 0; sh.is[I_HEAD] = 0; sh.sc[S_THETA] = 1.0; }
      __syncthreads();
    }
    // ================= search direction and line-search start =================
    double pdtd = 0.0, pgd = 0.0;
    for (int i = tid; i < D; i += LB_THREADS) {
      const double di = z[i] - x[i];
      d[i] = di;
      pdtd += di * di;
      pgd += g[i] * di;
    }
    const double dtd = lb_block_sum(pdtd, sh.red);
    const double gd0 = lb_block_sum(pgd, sh.red);
    // largest feasible step (first iteration: 1)
    double stpmx = LB_BIG;
    if (sh.is[I_ITER] == 0) stpmx = 1.0;
    else {
      // the sequential recurrence of the reference equals min over i of the per-variable limits (each update only lowers it)
      double loc = LB_BIG;
      for (int i = tid; i < D; i += LB_THREADS) {
        const double a1 = d[i];
        if (a1 < 0.0 && isfinite(l[i])) { const double a2 = l[i] - x[i]; loc = fmin(loc, (a2 >= 0.0) ? 0.0 : a2 / a1); }
        else if (a1 > 0.0 && isfinite(u[i])) { const double a2 = u[i] - x[i]; loc = fmin(loc, (a2 <= 0.0) ? 0.0 : a2 / a1); }
      }
      stpmx = -lb_block_max(-loc, sh.red);
    }
    int boxed = 1;
    for (int i = tid; i < D; i += LB_THREADS) if (!isfinite(l[i]) || !isfinite(u[i])) boxed = 0;
    const int all_boxed = (int)(lb_block_sum((double)(1 - boxed), sh.red) < 0.5);
    for (int i = tid; i < D; i += LB_THREADS) { t[i] = x[i]; gold[i] = g[i]; }
    __syncthreads();
    if (tid == 0) {
      const double dnorm = sqrt(dtd);
      const double stp = (sh.is[I_ITER] == 0 && !all_boxed) ? fmin(1.0 / dnorm, stpmx) : 1.0;
      sh.sc[S_DTD] = dtd; sh.sc[S_FOLD] = sh.sc[S_F]; sh.sc[S_GDOLD] = gd0;
      sh.sc[S_STP] = stp; sh.sc[S_STPMX] = stpmx;
      sh.is[I_IFUN] = 0; sh.is[I_IBACK] = 0;
      int fail = 0;
      if (gd0 >= 0.0) fail = 1;
      else {
        LbLs ls; lb_load_ls(ls, sh.sc, sh.is);
        if (lb_dcsrch(ls, sh.sc[S_F], gd0, true) != LS_FG) fail = 1;
        lb_store_ls(ls, sh.sc, sh.is);
      }
      sh.bi[4] = fail;
      if (!fail) { sh.is[I_IFUN] = 1; sh.is[I_IBACK] = 0; sh.is[I_PHASE] = LB_PH_LINESEARCH; }
    }
    __syncthreads();
    if (sh.bi[4]) { ls_failed = true; continue; }
    const double stp = sh.sc[S_STP];
    for (int i = tid; i < D; i += LB_THREADS) x[i] = (stp == 1.0) ? z[i] : stp * d[i] + t[i];
  }

  __syncthreads();
  if (tid < S_NSCALARS) ps[L.off_scal + tid] = sh.sc[tid];
  if (tid < I_NINTS) pi[tid] = sh.is[tid];
  for (int i = tid; i < LB_M * LB_M; i += LB_THREADS) { ps[L.off_ss + i] = sh.SS[i]; ps[L.off_sy + i] = sh.SY[i]; }
  if (tid == 0 && sh.is[I_TASK] == LB_FG && n_active != nullptr) atomicAdd(n_active, 1);
}

__global__ void lbfgsb_init_kernel(LbLayout L, const double* __restrict__ x0, const double* __restrict__ lower,
                                   const double* __restrict__ upper, double* __restrict__ X, double* __restrict__ state) {
  const int64_t b = blockIdx.x;
  double* ps = state + (size_t)b * L.per_problem;
  int* pi = reinterpret_cast<int*>(reinterpret_cast<char*>(state) + L.ints_offset_bytes) + (size_t)b * I_NINTS;
  for (size_t i = threadIdx.x; i < L.per_problem; i += blockDim.x) ps[i] = 0.0;
  __syncthreads();
  for (int i = threadIdx.x; i < L.D; i += blockDim.x)
    X[(size_t)b * L.D + i] = fmin(fmax(x0[(size_t)b * L.D + i], lower[i]), upper[i]);
  if (threadIdx.x < I_NINTS) pi[threadIdx.x] = 0;
  if (threadIdx.x == 0) ps[L.off_scal + S_THETA] = 1.0;
}

__global__ void lbfgsb_summary_kernel(LbLayout L, const double* __restrict__ state, double sign, double* __restrict__ f,
                                      int32_t* __restrict__ status) {
  const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (b >= L.N) return;
  const double* ps = state + (size_t)b * L.per_problem;
  const int* pi = reinterpret_cast<const int*>(reinterpret_cast<const char*>(state) + L.ints_offset_bytes) + (size_t)b * I_NINTS;
  f[b] = sign * ps[L.off_scal + S_F];
  status[4 * b + 0] = pi[I_TASK];
  status[4 * b + 1] = pi[I_MSG];
  status[4 * b + 2] = pi[I_ITER];
  status[4 * b + 3] = pi[I_NFEV];
}

}  // namespace mcacq

using namespace mcacq;

extern "C" size_t mcacq_lbfgsb_state_bytes(int64_t N, int D) {
  if (N < 0 || D <= 0) return 0;
  const LbLayout L = lb_layout(N, D);
  return L.ints_offset_bytes + (size_t)N * I_NINTS * sizeof(int) + 256;
}

extern "C" int mcacq_lbfgsb_init(int64_t N, int D, const double* x0, const double* lower, const double* upper, double* X,
                                 void* state, void* stream) {
  if (N < 0 || D <= 0 || !x0 || !lower || !upper || !X || !state) return MCACQ_EINVAL;
  if (N == 0) return 0;
  lbfgsb_init_kernel<<<(unsigned)N, 128, 0, (cudaStream_t)stream>>>(lb_layout(N, D), x0, lower, upper, X, (double*)state);
  count_launch();
  MCACQ_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" int mcacq_lbfgsb_step(int64_t N, int D, double* X, const double* f, const double* g, double sign,
                                 const double* lower, const double* upper, double factr, double pgtol, int maxiter, int maxfun,
                                 int maxls, void* state, int32_t* n_active, void* stream) {
  if (N < 0 || D <= 0 || !X || !f || !g || !lower || !upper || !state) return MCACQ_EINVAL;
  if (!(factr >= 0.0) || !(pgtol >= 0.0) || maxls <= 0) return MCACQ_EINVAL;
  if (N == 0) return 0;
  LbParams prm;
  prm.factr = factr; prm.pgtol = pgtol; prm.sign = sign; prm.maxiter = maxiter; prm.maxfun = maxfun; prm.maxls = maxls;
  if (n_active != nullptr) {
    cudaError_t e = cudaMemsetAsync(n_active, 0, sizeof(int32_t), (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
  }
  // W = [Y, theta S] (2 m x D) staged in shared memory per iteration when it fits next to the static working set
  const size_t w_bytes = (size_t)LB_2M * (size_t)((D + 1) & ~1) * sizeof(double);
  const int stage_w = w_bytes <= 160 * 1024 ? 1 : 0;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(lbfgsb_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  lbfgsb_step_kernel<<<(unsigned)N, LB_THREADS, stage_w ? w_bytes : 0, (cudaStream_t)stream>>>(
      lb_layout(N, D), X, f, g, lower, upper, (double*)state, prm, n_active, stage_w);
  count_launch();
  MCACQ_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" int mcacq_lbfgsb_summary(int64_t N, int D, const void* state, double sign, double* f, int32_t* status,
                                    void* stream) {
  if (N < 0 || D <= 0 || !state || !f || !status) return MCACQ_EINVAL;
  if (N == 0) return 0;
  lbfgsb_summary_kernel<<<(unsigned)((N + 127) / 128), 128, 0, (cudaStream_t)stream>>>(lb_layout(N, D), (const double*)state, sign,
                                                                                       f, status);
  count_launch();
  MCACQ_CUDA_CHECK_LAUNCH();
  return 0;
}
