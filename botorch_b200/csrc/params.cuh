// Parameter blocks shared by the stage kernels and the C-ABI orchestration.
#pragma once
#include <stdint.h>
namespace mcacq {
struct BlocksParams {
  int64_t b;
  int q, d, np, r;
  int kernel_id;
  double outputscale, mean_const, y_mean, y_std;
  const double* A;      // [b*q x np]
  const double* Kt;     // [b*q x np]   (nullptr in int8 mode: the mean comes from mean_part)
  const double* mean_part;  // [n_parts x b*q] per-column-tile partial sums of Kt * alpha (int8 mode)
  int n_parts;
  double* A_absmax;     // [b*q] max_k |A[i][k]| (optional; feeds the exponent bound of the backward slices)
  const double* alpha;  // [np]
  const double* U;      // [b*q x d]
  const double* A_base; // [r x np]
  const double* U_base; // [r x d]
  double* mean;         // [b*q]
  double* Sxx;          // [b x q x q]
  double* Sxb;          // [b x q x r_pitch]
  // baselines of more than 64 points are swept in chunks of 64 rows (one launch each): `r` is the chunk's row count, A_base /
  // U_base / Sxb point at the chunk's first row / column, `r_pitch` is the row pitch of Sxb (the whole baseline) and chunks
  // after the first only write their Sxb columns (`cross_only`)
  int r_pitch, cross_only;
  int32_t* counter;     // 4-byte scratch word for the tile counter of the baseline GEMM (r > 64); nullptr: chunked sweep
};

struct BlocksBwdParams {
  int64_t b;
  int q, d, np, r;
  int kernel_id;
  double outputscale, y_std;
  double* A;              // [b*q x np]  in: A, out: dA
  const double* A_base;   // [r x np]
  const double* U;        // [b*q x d]
  const double* U_base;   // [r x d]
  const double* gmean;    // [b*q]
  const double* gSxx;     // [b x q x q]
  const double* gSxb;     // [b x q x r]
  double* row_scale;      // [b*q]
  double* dU;             // [b*q x d]
  // int8 contraction mode: emit the G signed 8-bit slices of dA directly (no fp64 dA), scaled by a per-row bound
  int emit_slices, G;
  int8_t* slices;             // [G][b*q][np]
  double* slice_scale;        // [b*q]
  int32_t* slice_exp;         // [b*q]  scratch: exponent of the largest |dA[i][:]| (pass 1 -> pass 2)
  const double* A_absmax;     // [b*q]  max_k |A[i][k]| from the forward pass
  const double* Ab_absmax;    // [r]    max_k |A_base[j][k]|
  // r > 64: the baseline term gSxb A_base as a plain GEMM, computed beforehand (nullptr: run-time loop in the kernel)
  const double* T;            // [b*q x np]
  int skip_base_direct;       // 1: the direct K(X, X_base) terms of dU are added by baseline_direct_bwd_kernel afterwards
};

struct SRParams {
  int64_t b;
  int q, r, S, fat;
  double tau_relu, tau_max;
  const double* mean;    // [b*q]
  const double* Sxx;     // [b x q x q]
  const double* Sxb;     // [b x q x r]
  const double* L_base;  // [r x r]
  const double* Zt;      // [(r+q) x S]
  const double* best;    // [S]
  double* Bm;            // [b x q x r]
  double* Cm;            // [b x q x q]
  double* acq;           // [b]
  int32_t* info;         // [b]
  // round 2: affine objective, MC-mean utilities, smoothed outcome constraints (all optional; see mcacq_mc)
  double obj_w, obj_o;     // objective = obj_w * y + obj_o
  double util_param;       // modes 5 / 6: beta' / sqrt(pi / 2)
  const double* Zbar;      // [(r + q)] mean over the samples of every row of Zt (modes 5 / 6)
  int n_con, con_fat;
  double prior_var;      // outputscale * y_std^2 (0: no variance-collapse byte in the status word)
  int jitter_f32;        // jitter increments rounded to float32 first (see mcacq_mc.jitter_f32)
  double con_a[4], con_b[4], con_eta[4];
  // backward only
  const double* grad_acq;  // [b]
  double* gmean;           // [b*q]
  double* gSxx;            // [b x q x q]
  double* gSxb;            // [b x q x r]
};

}  // namespace mcacq
