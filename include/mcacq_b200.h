/*
 * mcacq_b200.h -- C ABI of the B200-native batched Monte-Carlo acquisition engine.
 *
 * Drop-in boundary: these entry points are what a BoTorch-side FFI for the
 * hot path `AcquisitionFunction.forward(X: b x q x d) -> b` and its gradient would bind.  The
 * reference's only native precedent is the pybind11 pair `forward` / `backward` of
 * botorch/csrc/logei_fused.cpp:376-379 wrapped by `_FusedLogAreas(autograd.Function)`
 * (botorch/acquisition/multi_objective/logei.py:107-169); the functions below play the same role
 * for qLogEI / qLogNEI on an exact SingleTaskGP:
 *
 *   mcacq_acq_forward      replaces  SampleReducingMCAcquisitionFunction.forward
 *                                    (botorch/acquisition/monte_carlo.py:268-305) with
 *                                    qLogNEI._get_samples_and_objectives (acquisition/logei.py:527-565),
 *                                    sample_cached_cholesky (utils/low_rank.py:84-172),
 *                                    _log_improvement (acquisition/logei.py:688-715),
 *                                    fatmax / logmeanexp (utils/safe_math.py:328-355, 213-225)
 *   mcacq_acq_backward     replaces  torch.autograd.grad(losses.sum(), X) (generation/gen.py:466-469)
 *   mcacq_posterior        replaces  BatchedMultiOutputGPyTorchModel.posterior (models/gpytorch.py:544-610)
 *                                    + Standardize.untransform_posterior (transforms/outcome.py:431-511)
 *   mcacq_scale_inputs     replaces  Normalize._transform (transforms/input.py:541-554) + x / lengthscale
 *   mcacq_cov_cross        replaces  gpytorch RBFKernel/MaternKernel(nu=2.5)/ScaleKernel forward
 *   mcacq_dgemm_tri        replaces  test_train_covar @ covar_cache (gpytorch exact_predictive_covar)
 *
 * Conventions: plain C, no exceptions, no ownership transfer.  Every pointer is a DEVICE pointer to
 * contiguous row-major fp64 unless it says "host".  The caller (torch) allocates all inputs, outputs
 * and the workspace; the library never allocates or frees device memory.  `stream` is a cudaStream_t
 * passed as void*.  Return value: 0 on success, a positive cudaError_t code on CUDA failure, or a
 * negative MCACQ_E* code on bad arguments.  All kernels are sm_100a only.
 */
#ifndef MCACQ_B200_H
#define MCACQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCACQ_KERNEL_RBF 0       /* o * exp(-0.5 * rho^2)                                  */
#define MCACQ_KERNEL_MATERN52 1  /* o * (1 + sqrt5 rho + 5/3 rho^2) exp(-sqrt5 rho)         */

#define MCACQ_TRI_UPPER 0 /* B[k][j] != 0 only for k <= j  (forward:  A = Kt * R)          */
#define MCACQ_TRI_LOWER 1 /* B[k][j] != 0 only for k >= j  (backward: dKt = dA * R^T)      */
#define MCACQ_TRI_DENSE 2

#define MCACQ_EINVAL (-1)  /* bad argument (null pointer, bad size, misaligned pitch)      */
#define MCACQ_ELIMIT (-2)  /* q, r, d or S outside the compiled limits                     */
#define MCACQ_EWORKSPACE (-3) /* workspace too small                                       */

#define MCACQ_MAX_Q 32
#define MCACQ_MAX_D 64
#define MCACQ_MAX_R 512 /* baselines beyond 64 points are swept in 64-row chunks; see mcacq_fused_supported */
#define MCACQ_MAX_SLICES 7 /* int8 contraction: at most 7 signed 8-bit slices (54-bit fixed-point operands) */

/* info[b] bits written by mcacq_acq_forward (psd_safe_cholesky semantics, max_tries = 6):   */
#define MCACQ_INFO_JITTER_MASK 0x7 /* number of jitter escalations applied (0 = none, 1 -> 1e-8, ... 6 -> 1e-3) */
#define MCACQ_INFO_NOT_PSD 0x8     /* still not PD after 6 tries  (reference: NotPSDError)   */
#define MCACQ_INFO_NONFINITE 0x10  /* NaN/Inf in the samples      (reference: NanError)      */
#define MCACQ_INFO_FLAG_MASK 0x1F  /* the bits above: 0 = nothing to report                    */
/* bits 8..15: conditioning of the q-batch, floor(-4 log2 rho) saturated at 255, rho = min_i C_ii^2 / Sxx_ii (the smallest
 * fraction of a point's posterior variance left after conditioning on the baseline draws and the preceding points).   */
#define MCACQ_INFO_COND_SHIFT 8
#define MCACQ_INFO_COND_MASK 0xFF00
/* bits 16..23: variance collapse of the q-batch, floor(-4 log2 min_i Sxx_ii / prior) saturated at 255 (prior = outputscale *
 * y_std^2): how many digits the difference prior - |a|^2 has lost.                                                     */
#define MCACQ_INFO_VAR_SHIFT 16
#define MCACQ_INFO_VAR_MASK 0xFF0000

/* Fitted-model operands (gpytorch DefaultPredictionStrategy caches + BoTorch transforms).   */
typedef struct {
  int32_t n;         /* train points                                                        */
  int32_t d;         /* input dim (<= MCACQ_MAX_D)                                          */
  int32_t np;        /* row pitch (in doubles) of every n-wide buffer; multiple of 16, >= n */
  int32_t kernel_id; /* MCACQ_KERNEL_*                                                      */
  double outputscale; /* ScaleKernel outputscale (1.0 if absent)                            */
  double mean_const;  /* ConstantMean constant                                              */
  double y_mean;      /* Standardize means (0 if no outcome transform)                      */
  double y_std;       /* Standardize stdvs (1 if no outcome transform)                      */
  const double* x_offset;    /* [d] Normalize offset      (zeros if no input transform)     */
  const double* x_coef;      /* [d] Normalize coefficient (ones  if no input transform)     */
  const double* lengthscale; /* [d] ARD lengthscales                                        */
  const double* U_train;     /* [n x d] ((X_train - offset) / coef) / lengthscale           */
  const double* alpha;       /* [np] mean_cache = (K + noise)^{-1} (y - c), zero padded     */
  const double* R;           /* [np x np] covar_cache = L^{-T}, upper triangular, zero padded */
  const double* Rt;          /* [np x np] transpose of R (lower triangular), zero padded    */
  /* Optional INT8-tensor-core contraction (csrc/ozaki_imma.cu).  contraction = 0: FP64 DMMA (mcacq_dgemm_tri);
   * contraction = 1: Ozaki split, g_fwd / g_bwd diagonals (<= MCACQ_MAX_SLICES; error ~2^-(8g-2) of row max x column max:
   * the host picks the smallest g whose build-time probe meets 1e-9 on the variance / 1e-7 on gradients).          */
  int32_t contraction;
  int32_t g_fwd;
  int32_t g_bwd;
  int32_t _pad;
  const int8_t* Rt_slices;   /* [g_fwd][np][np] row-scaled slices of R^T (row j = column j of R)  -- forward B operand */
  const double* Rt_scale;    /* [np]                                                                             */
  const int8_t* R_slices;    /* [g_bwd][np][np] row-scaled slices of R                            -- backward B operand */
  const double* R_scale;     /* [np]                                                                             */
} mcacq_model;

/* qLogNEI baseline operands (acquisition/logei.py:393-459); r == 0 for qLogEI.             */
typedef struct {
  int32_t r;             /* number of baseline points                                       */
  int32_t _pad;
  const double* U_base;  /* [r x d] scaled baseline inputs                                  */
  const double* A_base;  /* [r x np] K(X_base, X_train) R                                   */
  const double* L_base;  /* [r x r] lower Cholesky of the (untransformed) posterior cov of f(X_base) */
  const double* A_base_absmax; /* [r] max_k |A_base[j][k]| (optional; lets the int8 backward emit its slices fused) */
} mcacq_baseline;

/* Monte-Carlo operands.                                                                     */
typedef struct {
  int32_t S;           /* number of MC samples                                              */
  int32_t fat;         /* utility mode: 1 log_fatplus+fatmax+logmeanexp (qLogEI/qLogNEI), 0 log_softplus+smooth_amax+
                        * logmeanexp, 2 relu+amax+mean (qEI/qNEI), 3 identity+amax+mean (qSimpleRegret),
                        * 4 sigmoid((y-best)/tau_relu)+amax+mean (qProbabilityOfImprovement), 5 / 6 see util_param */
  double tau_relu;
  double tau_max;
  const double* Zt;    /* [(r + q) x S] base samples, TRANSPOSED (sample index contiguous)  */
  const double* best;  /* [S] per-sample incumbent (qLogEI: best_f repeated)                */
  /* Affine objective on the (single) outcome: obj = obj_weight * y + obj_offset -- `LinearMCObjective` on an m = 1 model
   * (botorch/acquisition/objective.py:318-358); 1, 0 = `IdentityMCObjective`.                                           */
  double obj_weight;
  double obj_offset;
  /* Utility modes whose per-sample value uses the MC mean mu_i = mean_s obj[s][i] (botorch/acquisition/monte_carlo.py):
   *   5  mu + util_param |obj - mu|, amax, mean   (qUpperConfidenceBound :896-906, util_param = sqrt(beta pi / 2);
   *                                                qLowerConfidenceBound :909-921 with the negative root)
   *   6  util_param |obj - mu|, amax, mean        (qPosteriorStandardDeviation :979-989, util_param = sqrt(pi / 2))
   * mu_i is evaluated in closed form, obj_weight (mean_i + sum_j coef_ij Zbar_j) + obj_offset, Zbar = row means of Zt.   */
  double util_param;
  const double* Zbar;  /* [(r + q)] (modes 5 / 6 only)                                                                   */
  /* Smoothed outcome constraints `c_k(y) = con_a[k] y + con_b[k] <= 0` (compute_smoothed_feasibility_indicator,
   * botorch/utils/objective.py:135-211): the per-sample utility is weighted by prod_k sigmoid(-c_k / con_eta[k]) (added as
   * log-sigmoid in the log modes 0 / 1); con_fat = 1 uses the fat-tailed `fatmoid` (safe_math.py:441-458).              */
  int32_t n_con;       /* 0 .. 4                                                                                         */
  int32_t con_fat;
  double con_a[4];
  double con_b[4];
  double con_eta[4];
  /* psd_safe_cholesky adds `(info > 0) * (jitter_new - jitter_prev)` to the diagonal: a bool tensor times a Python float,
   * i.e. a tensor of torch's DEFAULT dtype.  With the usual float32 default the increments 1e-8, 9e-8, 9e-7, ... are therefore
   * rounded to float32 before they reach the fp64 matrix (3e-11 absolute at the 1e-3 level).  1 = reproduce that (the
   * Python layer passes `torch.get_default_dtype() == torch.float32`), 0 = exact fp64 increments.                      */
  int32_t jitter_f32;
  int32_t _pad;
} mcacq_mc;

const char* mcacq_version(void);
int mcacq_num_sms(void);

/* u = ((x - offset) / coef) / lengthscale for `rows` points.  Replaces `Normalize._transform`
 * (botorch/models/transforms/input.py:541-554, applied by Model.transform_inputs, models/model.py:197-217) and the
 * `x.div(lengthscale)` of gpytorch's stationary kernels.                                      */
int mcacq_scale_inputs(const double* X, int64_t rows, int d, const double* x_offset, const double* x_coef,
                       const double* lengthscale, double* U, void* stream);

/* K[i][j] = k(U1[i], U2[j]) for i < m1, j < m2; columns m2..ldk-1 of each row are zero-filled.  Replaces the forward of
 * gpytorch's RBFKernel / MaternKernel(nu=2.5) / ScaleKernel as configured by botorch/models/utils/gpytorch_modules.py:100-133
 * and evaluated inside `ExactGP.__call__` (botorch/models/gpytorch.py:585).                      */
int mcacq_cov_cross(int kernel_id, double outputscale, const double* U1, int64_t m1, const double* U2, int m2,
                    int d, double* K, int64_t ldk, void* stream);

/* int8-contraction variant of mcacq_cov_cross: emits the G signed 8-bit slices [G][m1][ldk] of K (fixed exponent:
 * 0 < k <= outputscale < 2^fixed_exp), the per-512-column partial sums mean_part[ceil(ldk/512)][m1] of K * alpha, and
 * no fp64 matrix.  ldk must be a multiple of 16; columns m2..ldk-1 are zero.                                      */
int mcacq_cov_cross_sliced(int kernel_id, double outputscale, const double* U1, int64_t m1, const double* U2, int m2,
                           int d, int64_t ldk, const double* alpha, int G, int fixed_exp, int8_t* slices,
                           double* mean_part, void* stream);

/* dU1[i][:] = sum_j (W[i][j] + row_scale[i] * col_vec[j]) * d k(U1[i], U2[j]) / d U1[i]   (+= if accumulate; row_scale /
 * col_vec optional).  Replaces autograd through the kernel forward in `torch.autograd.grad(losses.sum(), X)`
 * (botorch/generation/gen.py:466-469).                                                         */
int mcacq_cov_cross_bwd(int kernel_id, double outputscale, const double* U1, int64_t m1, const double* U2, int m2,
                        int d, const double* W, int64_t ldw, const double* row_scale, const double* col_vec,
                        double* dU1, int accumulate, void* stream);

/* C[M x N] = A[M x K] * B[K x N], N == K == np (multiple of 16), fp64 DMMA, triangular-aware.  Replaces
 * `test_train_covar @ covar_cache` of gpytorch's exact_predictive_covar under fast_pred_var (the settings BoTorch
 * applies in botorch/models/utils/assorted.py:305-315) and its transpose product in the backward pass.
 * `tile_counter` is a 4-byte device scratch word (zeroed by the call).                      */
int mcacq_dgemm_tri(int tri_mode, int64_t M, int np, const double* A, const double* B, double* C,
                    int32_t* tile_counter, void* stream);

/* C[M x N] = A[M x K] * Bt[N x K]^T (both K-contiguous; lda, ldb, K even).  a_lower = 1: A is lower triangular
 * (k <= row), only that range is contracted.  Joint-posterior route (generation/sampling.py:89-155): SYRK
 * `A A^T` of gpytorch exact_predictive_covar and `L z` of MultivariateNormal.rsample.                          */
int mcacq_dgemm_nt(int a_lower, int64_t M, int N, int K, const double* A, int64_t lda, const double* Bt, int64_t ldb,
                   double* C, int64_t ldc, int32_t* tile_counter, void* stream);

/* Symmetric update C[N x N] := scale * (C - A A^T), A [N x K] (K-contiguous, lda and K even), C symmetric on entry: the
 * `test_test_covar - covar_correction_rhs^T covar_correction_rhs` of gpytorch exact_predictive_covar followed by
 * Standardize.untransform_posterior's s^2 (scale) for a joint posterior over N points.  Only the tiles on and below the
 * diagonal are contracted (half the flops of mcacq_dgemm_nt); every entry is stored at (i, j) and (j, i).             */
int mcacq_syrk_sub(int64_t N, int K, const double* A, int64_t lda, double* C, int64_t ldc, double scale,
                   int32_t* tile_counter, void* stream);

/* Y[N x S] = L[N x N] Z[S x N]^T for S <= 8 sample vectors, L lower triangular (only k <= row is read): the `root @
 * base_samples` of MultivariateNormal.rsample when MaxPosteriorSampling draws a handful of joint samples
 * (generation/sampling.py:89-155).  Memory-bound sweep over the triangle; any N, no alignment requirements.            */
int mcacq_lower_times_few(int64_t N, int S, const double* L, int64_t ldl, const double* Z, int64_t ldz, double* Y, int64_t ldy,
                          void* stream);

/* ---- FP64-accurate contraction on the INT8 tensor cores (Ozaki-style splitting; csrc/ozaki_imma.cu) -------------
 * Optional replacement of mcacq_dgemm_tri for `test_train_covar @ covar_cache`:
 *   mcacq_slice_rows    : X[rows x K] (fp64) -> G signed 8-bit slices [G][rows][Kp] (balanced radix-256 digits) + row scale;
 *   mcacq_ozaki_contract: C[M x N] = ea*fb * sum_{p+q<G} 256^-(p+q+2) A_p B_q^T  (tcgen05 kind::i8, exact int32
 *                         slice products, fp64 recombination); B slices are given as rows [N x K] (K contiguous).
 * tri_mode as in mcacq_dgemm_tri (which k-range of B^T is non-zero).                                              */
int mcacq_slice_rows(const double* X, int64_t rows, int K, int64_t ldx, int Kp, int G, int use_fixed_exp, int fixed_exp,
                     int8_t* slices, double* row_scale, void* stream);
int mcacq_ozaki_contract(int tri_mode, int64_t M, int N, int K, int G, const int8_t* A_slices, const double* row_scale,
                         const int8_t* B_slices, const double* col_scale, double* C, int64_t ldc, void* stream);

/* 1 if mcacq_acq_forward / _backward take this shape (q points per q-batch, r baseline points, S MC samples; mc_mean != 0:
 * the qUCB / qLCB / qPSTD utilities): the compiled limits AND the shared memory of the sample / reduce kernels, which grows
 * with q * r.  Host-only (no CUDA call).  Shapes it rejects take the generic route of
 * SampleReducingMCAcquisitionFunction (botorch/acquisition/monte_carlo.py:268-305) over mcacq_posterior.            */
int mcacq_fused_supported(int q, int r, int S, int mc_mean);

/* y[i] = f(x[i]) evaluated with the kernels' own FP64 elementary functions (csrc/fast_math.cuh; kind 0: exp, 1: log,
 * 2: log1p for x >= 0) -- what `torch.exp / log / log1p` are to the reference's safe_math (utils/safe_math.py:298-355).
 * Test hook: lets the parity suite bound their error (<= 2.5 ulp) on the device itself.                              */
int mcacq_fast_math_probe(int kind, const double* x, double* y, int64_t n, void* stream);

/* Workspace size of one forward(+backward) call.  `mcacq_workspace_bytes` is the bound over all contraction modes;
 * `mcacq_workspace_bytes_model` is exact for the model's mode (the FP64 DMMA mode carries no int8 slice buffers).   */
size_t mcacq_workspace_bytes(int64_t b, int q, int d, int np, int r);
size_t mcacq_workspace_bytes_model(const mcacq_model* model, int64_t b, int q, int r);

/* Posterior over b q-batches: mean [b x q], covar [b x q x q] on the original outcome scale. */
int mcacq_posterior(const mcacq_model* model, const double* X, int64_t b, int q, double* mean, double* covar,
                    void* workspace, size_t workspace_bytes, void* stream);

/* grad_X[b x q x d] = d( sum gmean * mean + sum gcovar * covar ) / dX, using the workspace left by the
 * matching mcacq_posterior call (consumed in place).  Replaces autograd through Model.posterior.      */
int mcacq_posterior_backward(const mcacq_model* model, const double* X, int64_t b, int q, const double* gmean,
                             const double* gcovar, double* grad_X, void* workspace, size_t workspace_bytes,
                             void* stream);

/* acq[b] = logmeanexp_S( fatmax_q( log_fatplus( y - best ) ) );  info[b] = Cholesky status.
 * The workspace must be kept intact between a forward and its backward.                      */
int mcacq_acq_forward(const mcacq_model* model, const mcacq_baseline* base, const mcacq_mc* mc, const double* X,
                      int64_t b, int q, double* acq, int32_t* info, void* workspace, size_t workspace_bytes,
                      void* stream);

/* The sample / reduce stage of mcacq_acq_forward on caller-provided posterior blocks -- mean [b x q], Sxx [b x q x q],
 * Sxb [b x q x r] (original outcome scale): replaces `sample_cached_cholesky` (botorch/utils/low_rank.py:84-172; r = 0:
 * `GPyTorchPosterior.rsample_from_base_samples`, posteriors/gpytorch.py:86-127) followed by `_sample_forward` and the
 * reductions.  Outputs: acq [b], info [b] (jitter level / NOT_PSD / NONFINITE bits), the factors Bm [b x q x r] and
 * Cm [b x q x q].  `base` supplies r and L_base only.                                                              */
int mcacq_sample_reduce_forward(const mcacq_baseline* base, const mcacq_mc* mc, const double* mean, const double* Sxx,
                                const double* Sxb, int64_t b, int q, double* acq, int32_t* info, double* Bm, double* Cm,
                                void* stream);

/* out3[0] = OR of the flag bits, out3[1] = largest conditioning byte, out3[2] = largest variance-collapse byte of the b status
 * words of mcacq_acq_forward (one small launch; the caller reads 12 bytes instead of scanning info[]).                  */
int mcacq_info_summary(const int32_t* info, int64_t b, int32_t* out3, void* stream);

/* grad_X[b x q x d] = d( sum_b grad_acq[b] * acq[b] ) / dX.                                  */
int mcacq_acq_backward(const mcacq_model* model, const mcacq_baseline* base, const mcacq_mc* mc, const double* X,
                       int64_t b, int q, const double* acq, const double* grad_acq, double* grad_X,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- qLogEHVI / qLogNEHVI inclusion-exclusion inner loop ---------------------------
 * CUDA counterpart of botorch/csrc/logei_fused.cpp: `fused_log_areas_forward` (:184-272) and
 * `fused_log_areas_backward` (:274-374), i.e. the pybind pair `forward` / `backward` (:376-379).
 *   obj_subsets [B x n_sub x isz x m], cell_lower / cell_upper [B_cells x nc x m] (batched_cells = 0: B_cells = 1,
 *   batched_cells = 1: B_cells = B)  ->  out [B x nc x n_sub];  backward: grad_out [B x nc x n_sub] ->
 *   grad_obj [B x n_sub x isz x m].  dtype: 0 = fp64, 1 = fp32.  isz <= 32, m <= 8 (MAX_I / MAX_M, :33-34).
 * lcl_workspace: B_cells * nc * m elements of scratch (log cell lengths).                                        */
int mcacq_log_areas_forward(const void* obj_subsets, const void* cell_lower, const void* cell_upper, int64_t B,
                            int n_sub, int isz, int m, int batched_cells, int nc, int dtype, double tau_relu,
                            double tau_max, void* out, void* lcl_workspace, void* stream);
int mcacq_log_areas_backward(const void* grad_out, const void* obj_subsets, const void* cell_lower,
                             const void* cell_upper, int64_t B, int n_sub, int isz, int m, int batched_cells, int nc,
                             int dtype, double tau_relu, double tau_max, void* grad_obj, void* lcl_workspace,
                             void* stream);

/* The whole inclusion-exclusion loop of `_compute_log_qehvi` (botorch/acquisition/multi_objective/logei.py:272-435, steps
 * 1-8) for B MC samples in one launch: obj [B x q x m] (objective samples of the q points), cell bounds [nc x m] ->
 * out[B] = logsumexp_cells logdiffexp(even-size subsets, odd-size subsets) of the log-areas; q <= 6, m <= 4, fp64.  Same
 * building blocks (log_fatplus, fatmin, upper-bound clamp) as mcacq_log_areas_*, each li[j][k] evaluated once per cell
 * instead of once per subset membership.  Backward: grad_out [B], out [B] (from the forward) -> grad_obj [B x q x m].      */
int mcacq_log_hvi_forward(const double* obj, const double* cell_lower, const double* cell_upper, int64_t B, int q, int m,
                          int nc, double tau_relu, double tau_max, double* out, double* lcl_workspace, void* stream);
int mcacq_log_hvi_backward(const double* grad_out, const double* out, const double* obj, const double* cell_lower,
                           const double* cell_upper, int64_t B, int q, int m, int nc, double tau_relu, double tau_max,
                           double* grad_obj, double* lcl_workspace, void* stream);

/* Scrambled-Sobol points straight on the device: out[k][j] = (shift[j] XOR_{b in gray(first_index + k)} sobolstate[j][b]) * 2^-30,
 * gray(i) = i ^ (i >> 1) -- the closed form of the sequence `torch.quasirandom.SobolEngine.draw` produces point by point
 * (botorch/utils/sampling.py:74-111 `draw_sobol_samples`, called by gen_batch_initial_conditions, optim/initializers.py:425-447).
 * `sobolstate` [dim][30] and `shift` [dim] are the engine's own scrambled direction numbers and digital shift (int64, device);
 * integer arithmetic, so the points are bit-identical to the host engine's (its float32-rounded first point is patched by the
 * caller).  Replaces the host draw + the H2D copy of `raw_samples x q x d` doubles per `optimize_acqf` call.           */
int mcacq_sobol_draw(const int64_t* sobolstate, const int64_t* shift, int dim, int64_t n, int64_t first_index, double* out,
                     void* stream);

/* ---- device-resident batched L-BFGS-B (csrc/lbfgsb.cu) ---------------------------------------
 * Replaces the host loop of botorch/optim/batched_lbfgs_b.py:365-634 (`fmin_l_bfgs_b_batched`: one scipy `setulb` state
 * machine per restart, all active restarts evaluated together) and the numpy <-> device hop of
 * botorch/generation/gen.py:423-485 (`f_np_wrapper`).  N independent problems of dimension D (= q * d), shared box
 * [lower, upper] (length D; +-inf = unbounded), history m = 10.  Reverse communication through device buffers:
 *   mcacq_lbfgsb_init   clamps x0 into X [N x D] and resets the state;
 *   mcacq_lbfgsb_step   consumes f [N] and g [N x D] evaluated at X (the optimiser MINIMISES sign * f, so sign = -1 feeds
 *                       acquisition values and their gradients directly), advances every still-active problem to its next
 *                       trial point, written into X, and counts the active problems into *n_active (device int, optional);
 *                       finished problems keep their final iterate in X;
 *   mcacq_lbfgsb_summary f [N] (sign removed) and status [N x 4] = (task, message, iterations, evaluations);
 *                       task: 0 active, 2 converged, 3 stopped (iteration / evaluation limit), 4 abnormal line search;
 *                       message: scipy's task word (401 pgtol, 402 factr, 502 maxfun, 504 maxiter).
 * factr / pgtol / maxiter / maxfun / maxls as in scipy (`ftol = factr * eps`).                                         */
size_t mcacq_lbfgsb_state_bytes(int64_t N, int D);
int mcacq_lbfgsb_init(int64_t N, int D, const double* x0, const double* lower, const double* upper, double* X, void* state,
                      void* stream);
int mcacq_lbfgsb_step(int64_t N, int D, double* X, const double* f, const double* g, double sign, const double* lower,
                      const double* upper, double factr, double pgtol, int maxiter, int maxfun, int maxls, void* state,
                      int32_t* n_active, void* stream);
int mcacq_lbfgsb_summary(int64_t N, int D, const void* state, double sign, double* f, int32_t* status, void* stream);

/* Number of kernels the last forward/backward call on this thread launched (for bench accounting). */
int mcacq_last_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MCACQ_B200_H */
