// C-ABI orchestration: workspace carving and the forward / backward / posterior pipelines.
//
// Forward  (per call, b q-batches, M = b*q rows):
//   U   = scale(X)                         scale_inputs_kernel
//   Kt  = k(U, U_train)            [M x np] cov_cross_kernel
//   A   = Kt R                     [M x np] dgemm_tri_kernel (upper)          <- dominant FLOPs
//   mean, Sxx, Sxb                          posterior_blocks_kernel
//   B, C, acq, info                         sample_reduce_fwd_kernel
// Backward:
//   gmean, gSxx, gSxb                       sample_reduce_bwd_kernel
//   dA (in place over A), row_scale, dU     posterior_blocks_bwd_kernel
//   dKt = dA R^T    (into the Kt buffer)    dgemm_tri_kernel (lower)          <- dominant FLOPs
//   dU += sum_k (dKt + s*gmean*alpha) dk/du cov_cross_bwd_kernel
//   dX  = dU / (coef * lengthscale)         unscale_grad_kernel
#include <cmath>
#include "common.cuh"

namespace mcacq {

// defined in blocks.cu / sample_reduce.cu / cov.cu
int unscale_grad(const double* dU, int64_t rows, int d, const double* coef, const double* ls, double* dX, cudaStream_t st);
int fill_value(double* p, int64_t n, double v, cudaStream_t st);
int cov_cross_bwd_split(int kernel_id, double outputscale, const double* U1, int64_t m1, const double* U2, int m2, int d,
                        double* W, int64_t ldw, const double* row_scale, const double* col_vec, double* dU1, int accumulate,
                        cudaStream_t st);

}  // namespace mcacq

#include "params.cuh"

namespace mcacq {

int posterior_blocks_fwd(const BlocksParams& p, cudaStream_t st);
int posterior_blocks_bwd(const BlocksBwdParams& p, cudaStream_t st);
int sample_reduce_fwd(const SRParams& p, cudaStream_t st);
int sample_reduce_bwd(const SRParams& p, cudaStream_t st);
int info_summary(const int32_t* info, int64_t b, int32_t* out, cudaStream_t st);
int transpose_panel(const double* src, int rows, int cols, double* dst, cudaStream_t st);
bool baseline_direct_fits(int q, int r);
int baseline_direct_bwd(const BlocksBwdParams& p, cudaStream_t st);
size_t sample_reduce_fwd_smem(int q, int r, int S);
size_t sample_reduce_bwd_smem(int q, int r, int S, bool mc_mean);

struct Workspace {
  double *U, *Kt, *A, *mean, *Sxx, *Sxb, *Bm, *Cm, *gmean, *gSxx, *gSxb, *row_scale, *dU, *slice_scale, *mean_part,
      *A_absmax;
  int32_t* slice_exp;
  int8_t* slices;
  int32_t* counter;
  double* AbT;   // [np x r] transposed baseline panel (r > 64: right operand of the backward baseline GEMM)
  size_t bytes;
};

static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

// int8_g: 0 = FP64 DMMA contraction (no slice buffers), otherwise the number of int8 slices of the M x np left operand
// (max of the forward and backward diagonals of the model).
static Workspace carve(void* base, int64_t b, int q, int d, int np, int r, int int8_g) {
  Workspace w;
  char* p = (char*)base;
  size_t off = 0;
  const int64_t M = b * q;
  auto take = [&](size_t nbytes) {
    char* out = p ? p + off : nullptr;
    off += al256(nbytes);
    return out;
  };
  w.counter = (int32_t*)take(256);
  w.U = (double*)take((size_t)M * d * 8);
  w.Kt = (double*)take((size_t)M * np * 8);
  w.A = (double*)take((size_t)M * np * 8);
  w.mean = (double*)take((size_t)M * 8);
  w.Sxx = (double*)take((size_t)M * q * 8);
  w.Sxb = (double*)take((size_t)M * (r > 0 ? r : 1) * 8);
  w.Bm = (double*)take((size_t)M * (r > 0 ? r : 1) * 8);
  w.Cm = (double*)take((size_t)M * q * 8);
  w.gmean = (double*)take((size_t)M * 8);
  w.gSxx = (double*)take((size_t)M * q * 8);
  w.gSxb = (double*)take((size_t)M * (r > 0 ? r : 1) * 8);
  w.row_scale = (double*)take((size_t)M * 8);
  w.dU = (double*)take((size_t)M * d * 8);
  // operands of the optional INT8 contraction: 6 slices of the M x np left operand + its row scales
  w.slice_scale = (double*)take((size_t)M * 8);
  w.A_absmax = (double*)take((size_t)M * 8);
  w.slice_exp = (int32_t*)take((size_t)M * 4);
  w.mean_part = (double*)take(int8_g ? (size_t)((np + 63) / 64) * M * 8 : 256);
  w.slices = (int8_t*)take(int8_g ? (size_t)int8_g * M * np : 256);
  w.AbT = (double*)take(r > 64 ? (size_t)r * np * 8 : 256);
  w.bytes = off;
  return w;
}

static inline int model_int8_g(const mcacq_model* m) {
  if (m->contraction != 1) return 0;
  return m->g_fwd > m->g_bwd ? m->g_fwd : m->g_bwd;
}

static int check_model(const mcacq_model* m) {
  if (!m) return MCACQ_EINVAL;
  if (m->n <= 0 || m->d <= 0 || m->np < m->n || (m->np % 16) != 0) return MCACQ_EINVAL;
  if (m->d > MCACQ_MAX_D) return MCACQ_ELIMIT;
  if (!m->x_offset || !m->x_coef || !m->lengthscale || !m->U_train || !m->alpha || !m->R || !m->Rt) return MCACQ_EINVAL;
  if (m->kernel_id != MCACQ_KERNEL_RBF && m->kernel_id != MCACQ_KERNEL_MATERN52) return MCACQ_EINVAL;
  if (m->contraction != 0 && m->contraction != 1) return MCACQ_EINVAL;
  if (m->contraction == 1) {
    if (!m->Rt_slices || !m->Rt_scale || !m->R_slices || !m->R_scale) return MCACQ_EINVAL;
    if (m->g_fwd < 1 || m->g_fwd > MCACQ_MAX_SLICES || m->g_bwd < 1 || m->g_bwd > MCACQ_MAX_SLICES) return MCACQ_EINVAL;
  }
  return 0;
}

// shared front end: U, Kt, A, blocks
static int run_posterior_stage(const mcacq_model* m, const mcacq_baseline* base, const double* X, int64_t b, int q,
                               Workspace& w, cudaStream_t st) {
  const int64_t M = b * q;
  const int r = base ? base->r : 0;
  int rc;
  if ((rc = mcacq_scale_inputs(X, M, m->d, m->x_offset, m->x_coef, m->lengthscale, w.U, st))) return rc;
  const bool int8 = (m->contraction == 1);
  if (int8) {
    // Kt in (0, outputscale]: one fixed exponent for all rows, 2^e > outputscale >= |Kt|.  The covariance kernel emits
    // the slices and the partial means directly; the fp64 Kt never exists in this mode.
    int e = 0;
    frexp(m->outputscale, &e);
    if ((rc = mcacq_cov_cross_sliced(m->kernel_id, m->outputscale, w.U, M, m->U_train, m->n, m->d, m->np, m->alpha,
                                     m->g_fwd, e, w.slices, w.mean_part, st)))
      return rc;
    // row scale of the fixed-exponent slices: 2^(e+2) for every row
    if ((rc = fill_value(w.slice_scale, M, ldexp(1.0, e + 2), st))) return rc;
    if ((rc = mcacq_ozaki_contract(MCACQ_TRI_UPPER, M, m->np, m->np, m->g_fwd, w.slices, w.slice_scale, m->Rt_slices,
                                   m->Rt_scale, w.A, m->np, st)))
      return rc;
  } else {
    if ((rc = mcacq_cov_cross(m->kernel_id, m->outputscale, w.U, M, m->U_train, m->n, m->d, w.Kt, m->np, st))) return rc;
    if ((rc = mcacq_dgemm_tri(MCACQ_TRI_UPPER, M, m->np, w.Kt, m->R, w.A, w.counter, st))) return rc;
  }
  BlocksParams bp;
  bp.b = b; bp.q = q; bp.d = m->d; bp.np = m->np; bp.r = r;
  bp.kernel_id = m->kernel_id; bp.outputscale = m->outputscale; bp.mean_const = m->mean_const;
  bp.y_mean = m->y_mean; bp.y_std = m->y_std;
  bp.A = w.A; bp.Kt = int8 ? nullptr : w.Kt; bp.alpha = m->alpha; bp.U = w.U;
  bp.mean_part = int8 ? w.mean_part : nullptr; bp.n_parts = ((m->np + 63) / 64 + 7) / 8;  // CT_PER_CTA = 8 column tiles per partial (cov.cu)
  bp.A_absmax = w.A_absmax;
  bp.A_base = r > 0 ? base->A_base : nullptr;
  bp.U_base = r > 0 ? base->U_base : nullptr;
  bp.mean = w.mean; bp.Sxx = w.Sxx; bp.Sxb = w.Sxb;
  bp.counter = w.counter;
  return posterior_blocks_fwd(bp, st);
}

}  // namespace mcacq

using namespace mcacq;

extern "C" const char* mcacq_version(void) { return "mcacq_b200 0.1.0 (sm_100a)"; }

extern "C" int mcacq_num_sms(void) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  return sms;
}

extern "C" int mcacq_last_launch_count(void) { return g_launch_count; }

extern "C" size_t mcacq_workspace_bytes(int64_t b, int q, int d, int np, int r) {
  if (b < 0 || q <= 0 || d <= 0 || np <= 0 || r < 0) return 0;
  return carve(nullptr, b, q, d, np, r, MCACQ_MAX_SLICES).bytes;  // worst case over the contraction modes
}

extern "C" size_t mcacq_workspace_bytes_model(const mcacq_model* model, int64_t b, int q, int r) {
  if (check_model(model) != 0 || b < 0 || q <= 0 || r < 0) return 0;
  return carve(nullptr, b, q, model->d, model->np, r, model_int8_g(model)).bytes;
}

extern "C" int mcacq_posterior(const mcacq_model* model, const double* X, int64_t b, int q, double* mean,
                               double* covar, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_model(model);
  if (rc) return rc;
  if (b < 0 || q <= 0) return MCACQ_EINVAL;
  if (b == 0) return 0;  // empty t-batch: nothing to launch, the (possibly null) pointers are never touched
  if (!X || !mean || !covar || !workspace || b < 0 || q <= 0) return MCACQ_EINVAL;
  if (q > MCACQ_MAX_Q) return MCACQ_ELIMIT;
  if (b == 0) return 0;
  g_launch_count = 0;
  Workspace w = carve(workspace, b, q, model->d, model->np, 0, model_int8_g(model));
  if (w.bytes > workspace_bytes) return MCACQ_EWORKSPACE;
  // write mean / covariance straight into the caller's buffers
  w.mean = mean;
  w.Sxx = covar;
  return run_posterior_stage(model, nullptr, X, b, q, w, (cudaStream_t)stream);
}

// shared tail of the backward pass: blocks_bwd -> dgemm (lower) -> cov bwd -> unscale
static int run_posterior_backward(const mcacq_model* model, const mcacq_baseline* base, int64_t b, int q,
                                  const double* gmean, const double* gSxx, const double* gSxb, Workspace& w,
                                  double* grad_X, cudaStream_t st) {
  const int r = base ? base->r : 0;
  const int64_t M = b * q;
  int rc;
  BlocksBwdParams bp;
  bp.b = b; bp.q = q; bp.d = model->d; bp.np = model->np; bp.r = r;
  bp.kernel_id = model->kernel_id; bp.outputscale = model->outputscale; bp.y_std = model->y_std;
  bp.A = w.A;
  bp.A_base = r > 0 ? base->A_base : nullptr;
  bp.U = w.U;
  bp.U_base = r > 0 ? base->U_base : nullptr;
  bp.gmean = gmean; bp.gSxx = gSxx; bp.gSxb = gSxb;
  bp.row_scale = w.row_scale; bp.dU = w.dU;
  // int8 mode: the kernel emits the slices of dA (scaled by a per-row bound) instead of the fp64 matrix
  const bool fuse_slices = (model->contraction == 1) && (r == 0 || base->A_base_absmax != nullptr);
  bp.emit_slices = fuse_slices ? 1 : 0; bp.G = model->g_bwd;
  bp.slices = w.slices; bp.slice_scale = w.slice_scale; bp.slice_exp = w.slice_exp; bp.A_absmax = w.A_absmax;
  bp.Ab_absmax = r > 0 ? base->A_base_absmax : nullptr;
  bp.T = nullptr;
  bp.skip_base_direct = 0;
  // large baselines: the baseline term of dA as a plain GEMM, T = gSxb A_base (into the Kt buffer, which is free until the
  // contraction below writes dKt there); odd r keeps the in-kernel loop (the 16-byte cp.async chunks need an even pitch)
  {
    const char* e = getenv("MCACQ_BIGR_GEMM");
    if (r > 64 && (r & 1) == 0 && !(e != nullptr && atoi(e) == 0)) {
      if ((rc = transpose_panel(base->A_base, r, model->np, w.AbT, st))) return rc;
      if ((rc = mcacq_dgemm_nt(0, M, model->np, r, gSxb, r, w.AbT, r, w.Kt, model->np, w.counter, st))) return rc;
      bp.T = w.Kt;
      bp.skip_base_direct = baseline_direct_fits(q, r) ? 1 : 0;
    }
  }
  if ((rc = posterior_blocks_bwd(bp, st))) return rc;
  if (bp.skip_base_direct && (rc = baseline_direct_bwd(bp, st))) return rc;
  if (model->contraction == 1) {
    if (!fuse_slices &&
        (rc = mcacq_slice_rows(w.A, M, model->np, model->np, model->np, model->g_bwd, 0, 0, w.slices, w.slice_scale, st)))
      return rc;
    if ((rc = mcacq_ozaki_contract(MCACQ_TRI_LOWER, M, model->np, model->np, model->g_bwd, w.slices, w.slice_scale,
                                   model->R_slices, model->R_scale, w.Kt, model->np, st)))
      return rc;
  } else {
    if ((rc = mcacq_dgemm_tri(MCACQ_TRI_LOWER, M, model->np, w.A, model->Rt, w.Kt, w.counter, st))) return rc;
  }
  if ((rc = cov_cross_bwd_split(model->kernel_id, model->outputscale, w.U, M, model->U_train, model->n, model->d, w.Kt,
                                model->np, w.row_scale, model->alpha, w.dU, /*accumulate=*/1, st)))
    return rc;
  return unscale_grad(w.dU, M, model->d, model->x_coef, model->lengthscale, grad_X, st);
}

extern "C" int mcacq_posterior_backward(const mcacq_model* model, const double* X, int64_t b, int q,
                                        const double* gmean, const double* gcovar, double* grad_X, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  int rc = check_model(model);
  if (rc) return rc;
  if (b < 0 || q <= 0) return MCACQ_EINVAL;
  if (b == 0) return 0;  // empty t-batch: nothing to launch, the (possibly null) pointers are never touched
  if (!X || !gmean || !gcovar || !grad_X || !workspace || b < 0 || q <= 0) return MCACQ_EINVAL;
  if (q > MCACQ_MAX_Q) return MCACQ_ELIMIT;
  if (b == 0) return 0;
  g_launch_count = 0;
  Workspace w = carve(workspace, b, q, model->d, model->np, 0, model_int8_g(model));
  if (w.bytes > workspace_bytes) return MCACQ_EWORKSPACE;
  return run_posterior_backward(model, nullptr, b, q, gmean, gcovar, nullptr, w, grad_X, (cudaStream_t)stream);
}

static int check_mc(const mcacq_mc* mc) {
  if (!mc || !mc->Zt || !mc->best || mc->S <= 0) return MCACQ_EINVAL;
  if (!(mc->tau_relu > 0.0) || !(mc->tau_max > 0.0)) return MCACQ_EINVAL;
  if (mc->fat < 0 || mc->fat > 6) return MCACQ_EINVAL;
  if ((mc->fat == 5 || mc->fat == 6) && !mc->Zbar) return MCACQ_EINVAL;
  if (mc->n_con < 0 || mc->n_con > 4) return MCACQ_ELIMIT;
  for (int k = 0; k < mc->n_con; k++) if (!(mc->con_eta[k] > 0.0)) return MCACQ_EINVAL;
  return 0;
}

static void fill_sr(SRParams& sp, const mcacq_baseline* base, const mcacq_mc* mc, int64_t b, int q, Workspace& w) {
  const int r = base ? base->r : 0;
  sp.b = b; sp.q = q; sp.r = r; sp.S = mc->S; sp.fat = mc->fat;
  sp.tau_relu = mc->tau_relu; sp.tau_max = mc->tau_max;
  sp.obj_w = mc->obj_weight; sp.obj_o = mc->obj_offset; sp.util_param = mc->util_param; sp.Zbar = mc->Zbar;
  sp.n_con = mc->n_con; sp.con_fat = mc->con_fat; sp.jitter_f32 = mc->jitter_f32; sp.prior_var = 0.0;
  for (int k = 0; k < 4; k++) { sp.con_a[k] = mc->con_a[k]; sp.con_b[k] = mc->con_b[k]; sp.con_eta[k] = mc->con_eta[k]; }
  sp.mean = w.mean; sp.Sxx = w.Sxx; sp.Sxb = w.Sxb;
  sp.L_base = r > 0 ? base->L_base : nullptr;
  sp.Zt = mc->Zt; sp.best = mc->best;
  sp.Bm = w.Bm; sp.Cm = w.Cm;
  sp.acq = nullptr; sp.info = nullptr; sp.grad_acq = nullptr;
  sp.gmean = w.gmean; sp.gSxx = w.gSxx; sp.gSxb = w.gSxb;
}

// y[i] = f(x[i]) with the kernels' own FP64 elementary functions (fast_math.cuh): kind 0 exp, 1 log, 2 log1p (x >= 0)
__global__ void fast_math_probe_kernel(int kind, const double* __restrict__ x, double* __restrict__ y, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = x[i];
  y[i] = kind == 0 ? fm_exp(v) : kind == 1 ? fm_log(v) : fm_log1p_nonneg(v);
}

extern "C" int mcacq_fast_math_probe(int kind, const double* x, double* y, int64_t n, void* stream) {
  if (kind < 0 || kind > 2 || !x || !y || n < 0) return MCACQ_EINVAL;
  if (n == 0) return 0;
  fast_math_probe_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(kind, x, y, n);
  return (int)cudaGetLastError();
}

extern "C" int mcacq_fused_supported(int q, int r, int S, int mc_mean) {
  if (q <= 0 || q > MCACQ_MAX_Q || r < 0 || r > MCACQ_MAX_R || S <= 0) return 0;
  const size_t limit = 200 * 1024;   // what the sample / reduce launchers accept
  return sample_reduce_fwd_smem(q, r, S) <= limit && sample_reduce_bwd_smem(q, r, S, mc_mean != 0) <= limit;
}

extern "C" int mcacq_acq_forward(const mcacq_model* model, const mcacq_baseline* base, const mcacq_mc* mc,
                                 const double* X, int64_t b, int q, double* acq, int32_t* info, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  int rc = check_model(model);
  if (rc) return rc;
  if ((rc = check_mc(mc))) return rc;
  if (b < 0 || q <= 0) return MCACQ_EINVAL;
  if (b == 0) return 0;  // empty t-batch: nothing to launch, the (possibly null) pointers are never touched
  if (!X || !acq || !info || !workspace || b < 0 || q <= 0) return MCACQ_EINVAL;
  if (q > MCACQ_MAX_Q) return MCACQ_ELIMIT;
  const int r = base ? base->r : 0;
  if (r < 0) return MCACQ_EINVAL;
  if (r > 0 && (!base->U_base || !base->A_base || !base->L_base)) return MCACQ_EINVAL;
  if (r > MCACQ_MAX_R) return MCACQ_ELIMIT;
  if (b == 0) return 0;
  g_launch_count = 0;
  Workspace w = carve(workspace, b, q, model->d, model->np, r, model_int8_g(model));
  if (w.bytes > workspace_bytes) return MCACQ_EWORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = run_posterior_stage(model, base, X, b, q, w, st))) return rc;
  SRParams sp;
  fill_sr(sp, base, mc, b, q, w);
  sp.acq = acq; sp.info = info;
  sp.prior_var = model->outputscale * model->y_std * model->y_std;
  return sample_reduce_fwd(sp, st);
}

extern "C" int mcacq_info_summary(const int32_t* info, int64_t b, int32_t* out3, void* stream) {
  if (!info || !out3 || b < 0) return MCACQ_EINVAL;
  return info_summary(info, b, out3, (cudaStream_t)stream);
}

extern "C" int mcacq_acq_backward(const mcacq_model* model, const mcacq_baseline* base, const mcacq_mc* mc,
                                  const double* X, int64_t b, int q, const double* acq, const double* grad_acq,
                                  double* grad_X, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_model(model);
  if (rc) return rc;
  if ((rc = check_mc(mc))) return rc;
  if (b < 0 || q <= 0) return MCACQ_EINVAL;
  if (b == 0) return 0;  // empty t-batch: nothing to launch, the (possibly null) pointers are never touched
  if (!X || !acq || !grad_acq || !grad_X || !workspace || b < 0 || q <= 0) return MCACQ_EINVAL;
  if (q > MCACQ_MAX_Q) return MCACQ_ELIMIT;
  const int r = base ? base->r : 0;
  if (r < 0 || r > MCACQ_MAX_R) return r < 0 ? MCACQ_EINVAL : MCACQ_ELIMIT;
  if (b == 0) return 0;
  g_launch_count = 0;
  Workspace w = carve(workspace, b, q, model->d, model->np, r, model_int8_g(model));
  if (w.bytes > workspace_bytes) return MCACQ_EWORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  SRParams sp;
  fill_sr(sp, base, mc, b, q, w);
  sp.acq = const_cast<double*>(acq);
  sp.grad_acq = grad_acq;
  if ((rc = sample_reduce_bwd(sp, st))) return rc;

  return run_posterior_backward(model, base, b, q, w.gmean, w.gSxx, w.gSxb, w, grad_X, st);
}

// Sample / reduce stage alone on caller-provided posterior blocks (mean, Sxx, Sxb on the original outcome scale): what
// `sample_cached_cholesky` + `_sample_forward` + the q- and sample reductions compute once the posterior is known.  Lets
// tests drive `psd_safe_cholesky`'s jitter ladder with covariance blocks of a KNOWN escalation level.
extern "C" int mcacq_sample_reduce_forward(const mcacq_baseline* base, const mcacq_mc* mc, const double* mean, const double* Sxx,
                                           const double* Sxb, int64_t b, int q, double* acq, int32_t* info, double* Bm,
                                           double* Cm, void* stream) {
  int rc = check_mc(mc);
  if (rc) return rc;
  if (b < 0 || q <= 0) return MCACQ_EINVAL;
  if (b == 0) return 0;
  const int r = base ? base->r : 0;
  if (r < 0 || !mean || !Sxx || !acq || !info || !Cm || (r > 0 && (!Sxb || !Bm || !base->L_base))) return MCACQ_EINVAL;
  if (q > MCACQ_MAX_Q || r > MCACQ_MAX_R) return MCACQ_ELIMIT;
  g_launch_count = 0;
  SRParams sp;
  Workspace w = {};
  w.mean = const_cast<double*>(mean); w.Sxx = const_cast<double*>(Sxx); w.Sxb = const_cast<double*>(Sxb);
  w.Bm = Bm; w.Cm = Cm;
  fill_sr(sp, base, mc, b, q, w);
  sp.acq = acq; sp.info = info;
  return sample_reduce_fwd(sp, (cudaStream_t)stream);
}
