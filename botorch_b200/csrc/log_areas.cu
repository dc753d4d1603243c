// Fused log-area kernels for the qLogEHVI / qLogNEHVI inclusion-exclusion inner loop.
//
// CUDA counterpart of the reference's only native kernel, botorch/csrc/logei_fused.cpp
// (`fused_log_areas_forward` :184-272, `fused_log_areas_backward` :274-374), which is CPU-only
// (gate: acquisition/multi_objective/logei.py:320-328).  Same I/O contract:
//   obj_subsets (B, n_sub, i, m), cell_lower / cell_upper (B_cells, nc, m) with B_cells in {1, B}
//   -> log-areas (B, nc, n_sub);   backward: grad (B, nc, n_sub) -> grad_obj (B, n_sub, i, m).
// Arithmetic follows the reference line by line: log_fatplus with the "safe" softplus (:39-48, :85-112), the
// fatmin smooth minimum with its n == 1 and < -1e29 branches (:116-177), the pairwise fatmin with the log cell
// length (upper bound clamped to 1e10 / 1e8, :218-219, :229-238), and the sum over objectives in index order.
//
// Mapping: objectives are independent inside one (b, subset, cell) item, so
//   forward : one thread per output element (b, c, s) -- s fastest, stores fully coalesced;
//   backward: one thread per (b, s, k); it walks the cells sequentially (same accumulation order as the
//             reference's `g_buf`), keeping the i partial gradients of objective k in registers.
#include "common.cuh"

namespace mcacq {

template <typename T> struct LAConst;
template <> struct LAConst<double> { static __device__ __forceinline__ double clamp() { return 1e10; } };
template <> struct LAConst<float> { static __device__ __forceinline__ float clamp() { return 1e8f; } };

template <typename T>
__device__ __forceinline__ T la_softplus(T y) {
  if (y > T(20)) return y;
  if (y < T(-20)) return exp(y);
  return log1p(exp(y));
}

template <typename T>
__device__ __forceinline__ T la_sigmoid(T y) {
  if (y >= T(0)) { T e = exp(-y); return T(1) / (T(1) + e); }
  T e = exp(y);
  return e / (T(1) + e);
}

template <typename T, bool GRAD>
__device__ __forceinline__ T la_log_fatplus(T x, T tau, T inv_tau, T& grad) {
  const T y = x * inv_tau;
  const T cy = T(1) / (T(1) + y * y);
  const T f = la_softplus(y) + T(0.1) * cy;
  const T tf = tau * f;
  if (!(tf > T(0))) { if (GRAD) grad = T(0); return T(-1e30); }
  if (GRAD) grad = (la_sigmoid(y) - T(0.2) * y * cy * cy) / tf;
  return log(tf);
}

// fatmin over n values (n <= NMAX); optional gradient weights.
template <typename T, int NMAX, bool GRAD>
__device__ __forceinline__ T la_fatmin(const T (&x)[NMAX], int n, T tau, T inv_tau, T (&gw)[NMAX]) {
  if (n == 1) { if (GRAD) gw[0] = T(1); return x[0]; }
  T mn = x[0];
  int ami = 0;
#pragma unroll
  for (int j = 1; j < NMAX; j++) if (j < n && x[j] < mn) { mn = x[j]; ami = j; }
  if (mn < T(-1e29)) {
    if (GRAD) {
#pragma unroll
      for (int j = 0; j < NMAX; j++) if (j < n) gw[j] = (j == ami) ? T(1) : T(0);
    }
    return mn;
  }
  T S = T(0), S_pd = T(0);
  T pd[NMAX];
#pragma unroll
  for (int j = 0; j < NMAX; j++) if (j < n) {
    const T z = (x[j] - mn) * inv_tau;
    const T d = T(2) + T(2) * z + z * z;
    S += T(2) / d;
    if (GRAD) { pd[j] = T(-2) * (T(2) + T(2) * z) / (d * d); S_pd += pd[j]; }
  }
  if (GRAD) {
#pragma unroll
    for (int j = 0; j < NMAX; j++) if (j < n) gw[j] = (j == ami) ? T(1) + (S_pd - T(-1)) / S : -pd[j] / S;
  }
  return mn - tau * log(S);
}

template <typename T>
__global__ void log_cell_length_kernel(const T* __restrict__ cl, const T* __restrict__ cu, int64_t total, T* __restrict__ lcl) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < total) lcl[i] = log(min(cu[i], LAConst<T>::clamp()) - cl[i]);
}

template <typename T, int IMAX>
__global__ void __launch_bounds__(256)
log_areas_fwd_kernel(const T* __restrict__ obj, const T* __restrict__ cl, const T* __restrict__ lcl, int64_t B,
                     int n_sub, int isz, int m, int batched_cells, int nc, T tau_relu, T tau_max, T* __restrict__ out) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t total = B * nc * n_sub;
  if (idx >= total) return;
  const int s = (int)(idx % n_sub);
  const int c = (int)((idx / n_sub) % nc);
  const int64_t b = idx / ((int64_t)n_sub * nc);
  const int64_t bc = batched_cells ? b : 0;
  const T inv_tr = T(1) / tau_relu, inv_tm = T(1) / tau_max;
  const T* o = obj + ((b * n_sub + s) * isz) * m;
  const T* lo = cl + (bc * nc + c) * m;
  const T* ll = lcl + (bc * nc + c) * m;
  T area = T(0);
  for (int k = 0; k < m; k++) {
    T li[IMAX], gdummy[IMAX];
    const T lok = lo[k];
#pragma unroll
    for (int j = 0; j < IMAX; j++) {
      T gd;
      li[j] = (j < isz) ? la_log_fatplus<T, false>(o[j * m + k] - lok, tau_relu, inv_tr, gd) : T(0);
    }
    const T lim = la_fatmin<T, IMAX, false>(li, isz, tau_max, inv_tm, gdummy);
    T pair[2] = {lim, ll[k]}, pg[2];
    area += la_fatmin<T, 2, false>(pair, 2, tau_max, inv_tm, pg);
  }
  out[idx] = area;
}

template <typename T, int IMAX>
__global__ void __launch_bounds__(256)
log_areas_bwd_kernel(const T* __restrict__ go, const T* __restrict__ obj, const T* __restrict__ cl,
                     const T* __restrict__ lcl, int64_t B, int n_sub, int isz, int m, int batched_cells, int nc,
                     T tau_relu, T tau_max, T* __restrict__ gobj) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t total = B * n_sub * m;
  if (idx >= total) return;
  const int k = (int)(idx % m);
  const int s = (int)((idx / m) % n_sub);
  const int64_t b = idx / ((int64_t)m * n_sub);
  const int64_t bc = batched_cells ? b : 0;
  const T inv_tr = T(1) / tau_relu, inv_tm = T(1) / tau_max;
  const T* o = obj + ((b * n_sub + s) * isz) * m + k;
  T ov[IMAX], gacc[IMAX];
#pragma unroll
  for (int j = 0; j < IMAX; j++) { ov[j] = (j < isz) ? o[j * m] : T(0); gacc[j] = T(0); }
  for (int c = 0; c < nc; c++) {
    const T g_out = go[(b * nc + c) * n_sub + s];
    const T lok = cl[(bc * nc + c) * m + k];
    T li[IMAX], lg[IMAX], fmg[IMAX];
#pragma unroll
    for (int j = 0; j < IMAX; j++) {
      if (j < isz) li[j] = la_log_fatplus<T, true>(ov[j] - lok, tau_relu, inv_tr, lg[j]);
      else { li[j] = T(0); lg[j] = T(0); }
    }
    const T lim = la_fatmin<T, IMAX, true>(li, isz, tau_max, inv_tm, fmg);
    T pair[2] = {lim, lcl[(bc * nc + c) * m + k]}, pg[2];
    la_fatmin<T, 2, true>(pair, 2, tau_max, inv_tm, pg);
    const T g_lim = g_out * pg[0];
#pragma unroll
    for (int j = 0; j < IMAX; j++) if (j < isz) gacc[j] += g_lim * fmg[j] * lg[j];
  }
  T* g = gobj + ((b * n_sub + s) * isz) * m + k;
#pragma unroll
  for (int j = 0; j < IMAX; j++) if (j < isz) g[j * m] = gacc[j];
}

template <typename T>
static int log_areas_run(int backward, const T* go, const T* obj, const T* cl, const T* cu, int64_t B, int n_sub, int isz,
                         int m, int batched_cells, int nc, double tau_relu, double tau_max, T* out, T* lcl,
                         cudaStream_t st) {
  const int64_t ncell = (batched_cells ? B : 1) * (int64_t)nc * m;
  log_cell_length_kernel<T><<<(unsigned)((ncell + 255) / 256), 256, 0, st>>>(cl, cu, ncell, lcl);
  count_launch();
  const int64_t total = backward ? B * n_sub * m : B * nc * n_sub;
  const unsigned blocks = (unsigned)((total + 255) / 256);
#define LA_CASE(IM)                                                                                                 \
  if (isz <= IM) {                                                                                                  \
    if (backward)                                                                                                   \
      log_areas_bwd_kernel<T, IM><<<blocks, 256, 0, st>>>(go, obj, cl, lcl, B, n_sub, isz, m, batched_cells, nc,     \
                                                           (T)tau_relu, (T)tau_max, out);                           \
    else                                                                                                            \
      log_areas_fwd_kernel<T, IM><<<blocks, 256, 0, st>>>(obj, cl, lcl, B, n_sub, isz, m, batched_cells, nc,         \
                                                           (T)tau_relu, (T)tau_max, out);                           \
    count_launch();                                                                                                 \
    MCACQ_CUDA_CHECK_LAUNCH();                                                                                      \
    return 0;                                                                                                       \
  }
  LA_CASE(2) LA_CASE(4) LA_CASE(8) LA_CASE(16) LA_CASE(32)
#undef LA_CASE
  return MCACQ_ELIMIT;
}

}  // namespace mcacq

static int la_check(const void* obj, const void* cl, const void* cu, const void* out, const void* lcl, int64_t B,
                    int n_sub, int isz, int m, int nc, int dtype, double tau_relu, double tau_max) {
  if (!obj || !cl || !cu || !out || !lcl || B < 0 || n_sub <= 0 || isz <= 0 || m <= 0 || nc <= 0) return MCACQ_EINVAL;
  if (dtype != 0 && dtype != 1) return MCACQ_EINVAL;
  if (!(tau_relu > 0.0) || !(tau_max > 0.0)) return MCACQ_EINVAL;
  if (isz > 32 || m > 8) return MCACQ_ELIMIT;  // MAX_I / MAX_M of the reference (logei_fused.cpp:33-34)
  return 0;
}

extern "C" int mcacq_log_areas_forward(const void* obj_subsets, const void* cell_lower, const void* cell_upper,
                                       int64_t B, int n_sub, int isz, int m, int batched_cells, int nc, int dtype,
                                       double tau_relu, double tau_max, void* out, void* lcl_workspace, void* stream) {
  using namespace mcacq;
  int rc = la_check(obj_subsets, cell_lower, cell_upper, out, lcl_workspace, B, n_sub, isz, m, nc, dtype, tau_relu, tau_max);
  if (rc) return rc;
  if (B == 0) return 0;
  g_launch_count = 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == 0)
    return log_areas_run<double>(0, nullptr, (const double*)obj_subsets, (const double*)cell_lower, (const double*)cell_upper,
                                 B, n_sub, isz, m, batched_cells, nc, tau_relu, tau_max, (double*)out, (double*)lcl_workspace, st);
  return log_areas_run<float>(0, nullptr, (const float*)obj_subsets, (const float*)cell_lower, (const float*)cell_upper, B,
                              n_sub, isz, m, batched_cells, nc, tau_relu, tau_max, (float*)out, (float*)lcl_workspace, st);
}

extern "C" int mcacq_log_areas_backward(const void* grad_out, const void* obj_subsets, const void* cell_lower,
                                        const void* cell_upper, int64_t B, int n_sub, int isz, int m, int batched_cells,
                                        int nc, int dtype, double tau_relu, double tau_max, void* grad_obj,
                                        void* lcl_workspace, void* stream) {
  using namespace mcacq;
  if (!grad_out) return MCACQ_EINVAL;
  int rc = la_check(obj_subsets, cell_lower, cell_upper, grad_obj, lcl_workspace, B, n_sub, isz, m, nc, dtype, tau_relu, tau_max);
  if (rc) return rc;
  if (B == 0) return 0;
  g_launch_count = 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == 0)
    return log_areas_run<double>(1, (const double*)grad_out, (const double*)obj_subsets, (const double*)cell_lower,
                                 (const double*)cell_upper, B, n_sub, isz, m, batched_cells, nc, tau_relu, tau_max,
                                 (double*)grad_obj, (double*)lcl_workspace, st);
  return log_areas_run<float>(1, (const float*)grad_out, (const float*)obj_subsets, (const float*)cell_lower,
                              (const float*)cell_upper, B, n_sub, isz, m, batched_cells, nc, tau_relu, tau_max,
                              (float*)grad_obj, (float*)lcl_workspace, st);
}
