"""Device-resident exact-GP prediction caches and the C-ABI model descriptor.

B200-native counterpart of gpytorch's `DefaultPredictionStrategy` as BoTorch configures it
(botorch/__init__.py:52-57, botorch/models/utils/assorted.py:305-315): the train Cholesky factor `L`,
`mean_cache = (K + noise)^{-1} (y - c)` and `covar_cache = L^{-T}` are built once per fitted model
 and stay resident in HBM; every posterior / acquisition call then only
runs the hand-written kernels of `csrc/`.  The one-off factorisation uses torch's cuSOLVER Cholesky and
triangular solve (a plain library call on an n x n matrix, off the per-evaluation path).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import warnings

import torch
from torch import Tensor

from .. import _lib, settings
from ..exceptions.errors import NanError, NotPSDError
from ..exceptions.warnings import NumericalWarning


def psd_safe_cholesky(A: Tensor, max_tries: int = 6, jitter: float | None = None) -> Tensor:
    """`linear_operator.utils.cholesky.psd_safe_cholesky` semantics (jitter only on failing batch
    elements, 1e-8 * 10^i in fp64), evaluated with torch on the tensor's device."""
    L, info = torch.linalg.cholesky_ex(A)
    if not torch.any(info):
        return L
    if torch.isnan(A).any():
        raise NanError(f"cholesky: {int(torch.isnan(A).sum())} of {A.numel()} elements are NaN.")
    if jitter is None:
        jitter = 1e-6 if A.dtype == torch.float32 else 1e-8
    Aprime = A.clone()
    jitter_prev = 0.0
    for i in range(max_tries):
        jitter_new = jitter * (10**i)
        diag_add = ((info > 0) * (jitter_new - jitter_prev)).unsqueeze(-1).expand(*Aprime.shape[:-1])
        Aprime.diagonal(dim1=-1, dim2=-2).add_(diag_add)
        jitter_prev = jitter_new
        warnings.warn(f"A not p.d., added jitter of {jitter_new:.1e} to the diagonal", NumericalWarning)
        L, info = torch.linalg.cholesky_ex(Aprime)
        if not torch.any(info):
            return L
    raise NotPSDError(f"Matrix not positive definite after repeatedly adding jitter up to {jitter_new:.1e}.")


class DevicePredictionStrategy:
    """Holds the fitted-state operands of `mcacq_model` on one CUDA device."""

    def __init__(
        self,
        train_X_transformed: Tensor,  # n x d, already through the input transform
        train_Y_standardized: Tensor,  # n (standardised targets)
        lengthscale: Tensor,  # d
        noise: Tensor,  # scalar or n
        kernel_id: int,
        outputscale: float,
        mean_const: float,
        x_offset: Tensor,  # d
        x_coef: Tensor,  # d
        y_mean: float,
        y_std: float,
        contraction: str = "dmma",
    ) -> None:
        dev = train_X_transformed.device
        if dev.type != "cuda":
            raise _lib.McacqError("DevicePredictionStrategy requires CUDA tensors; botorch_b200 has no CPU path.")
        f64 = dict(device=dev, dtype=torch.float64)
        self.device = dev
        self.n, self.d = train_X_transformed.shape
        self.np = _lib.round_up(self.n, 16)
        self.kernel_id, self.outputscale, self.mean_const = kernel_id, float(outputscale), float(mean_const)
        self.y_mean, self.y_std = float(y_mean), float(y_std)
        self.x_offset = x_offset.to(**f64).contiguous()
        self.x_coef = x_coef.to(**f64).contiguous()
        self.lengthscale = lengthscale.to(**f64).reshape(-1).contiguous()
        L = _lib.lib()
        st = _lib.stream_ptr()
        n, d, npad = self.n, self.d, self.np
        # scaled train inputs: the train inputs are already input-transformed, so only divide by l
        zeros, ones = torch.zeros(d, **f64), torch.ones(d, **f64)
        Xt = train_X_transformed.to(**f64).contiguous()
        self.U_train = torch.empty(n, d, **f64)
        _lib.check(L.mcacq_scale_inputs(Xt.data_ptr(), n, d, zeros.data_ptr(), ones.data_ptr(),
                                        self.lengthscale.data_ptr(), self.U_train.data_ptr(), st), "scale_inputs")
        # K(X_train, X_train) with the hand-written covariance kernel, + noise on the diagonal
        Ktt = torch.empty(n, npad, **f64)
        _lib.check(L.mcacq_cov_cross(kernel_id, self.outputscale, self.U_train.data_ptr(), n, self.U_train.data_ptr(),
                                     n, d, Ktt.data_ptr(), npad, st), "cov_cross(train, train)")
        Khat = Ktt[:, :n].contiguous()
        noise = noise.to(**f64).reshape(-1)
        Khat.diagonal().add_(noise.expand(n))
        chol = psd_safe_cholesky(Khat)
        y = train_Y_standardized.to(**f64).reshape(-1)
        alpha = torch.cholesky_solve((y - self.mean_const).unsqueeze(-1), chol).squeeze(-1)
        eye = torch.eye(n, **f64)
        Linv = torch.linalg.solve_triangular(chol, eye, upper=False)
        del eye, Khat, Ktt
        self.alpha = torch.zeros(npad, **f64)
        self.alpha[:n] = alpha
        self.R = torch.zeros(npad, npad, **f64)
        self.R[:n, :n] = Linv.mT
        self.Rt = torch.zeros(npad, npad, **f64)
        self.Rt[:n, :n] = Linv
        del Linv
        self.train_chol = chol
        # INT8 tensor-core contraction: signed 8-bit row-scaled slices of R^T (forward) and R (backward)
        self.contraction = contraction
        if contraction not in ("dmma", "int8"):
            raise ValueError("contraction must be 'dmma' or 'int8'")
        self.Rt_slices = self.Rt_scale = self.R_slices = self.R_scale = None
        self.g_fwd, self.g_bwd = self.G_FWD_LADDER[0], self.G_BWD_LADDER[0]
        self.desc = _lib.Model(
            n=n, d=d, np=npad, kernel_id=kernel_id, outputscale=self.outputscale, mean_const=self.mean_const,
            y_mean=self.y_mean, y_std=self.y_std, x_offset=self.x_offset.data_ptr(), x_coef=self.x_coef.data_ptr(),
            lengthscale=self.lengthscale.data_ptr(), U_train=self.U_train.data_ptr(), alpha=self.alpha.data_ptr(),
            R=self.R.data_ptr(), Rt=self.Rt.data_ptr(),
            contraction=0, g_fwd=self.g_fwd, g_bwd=self.g_bwd, _pad=0,
            Rt_slices=None, Rt_scale=None, R_slices=None, R_scale=None,
        )
        self.int8_probe_error = self.int8_probe_grad_error = None
        lim = settings.int8_cond_limit.value()
        self.int8_cond_limit = self.INT8_COND_LIMIT if lim == "auto" else lim
        self._fp64_view = None
        self._max_view = None
        self.int8_var_ratio_limit, self.int8_var_byte_limit = None, None
        if contraction == "int8":
            self._select_int8(Xt)
            if lim is None:   # settings.int8_cond_limit(None) switches every run-time re-routing off (developer tools)
                self.int8_var_ratio_limit, self.int8_var_byte_limit = None, None

    def max_slices_view(self) -> "DevicePredictionStrategy":
        """The same fitted state with the largest exact slice counts (7 / 7 when n allows): 2^-54 of the prior, i.e. FP64-level
        accuracy without the run-time re-routing.  For callers that cannot look at the status words between launches (the
        CUDA-graph rounds of the device-resident optimiser); those calls are a few hundred rows, where slices cost nothing."""
        if self.contraction != "int8":
            return self
        if getattr(self, "_max_view", None) is None:
            import copy

            G = max(g for g in self.G_BWD_LADDER if self._int8_exact(g))
            v = copy.copy(self)
            v.desc = _lib.Model.from_buffer_copy(self.desc)
            v._fp64_view = None
            v._max_view = v
            v.Rt_slices, v.Rt_scale = (self.Rt_slices, self.Rt_scale) if self.g_fwd == G else self._slice_rows(self.Rt, G)
            v.R_slices, v.R_scale = (self.R_slices, self.R_scale) if self.g_bwd == G else self._slice_rows(self.R, G)
            v.g_fwd = v.g_bwd = G
            v.desc.g_fwd = v.desc.g_bwd = G
            v.desc.Rt_slices, v.desc.Rt_scale = v.Rt_slices.data_ptr(), v.Rt_scale.data_ptr()
            v.desc.R_slices, v.desc.R_scale = v.R_slices.data_ptr(), v.R_scale.data_ptr()
            v.int8_var_ratio_limit, v.int8_var_byte_limit = None, None
            self._max_view = v
        return self._max_view

    def fp64_view(self) -> "DevicePredictionStrategy":
        """The same fitted state with the FP64 DMMA contraction: where the int8 mode sends ill-conditioned q-batches."""
        if self.contraction != "int8":
            return self
        if self._fp64_view is None:
            import copy

            v = copy.copy(self)
            v.desc = _lib.Model.from_buffer_copy(self.desc)
            v.desc.contraction = 0
            v.contraction = "dmma"
            v._fp64_view = None
            self._fp64_view = v
        return self._fp64_view

    # The slices are FIXED-point relative to each row's largest entry, so the int8 contraction reproduces the posterior
    # variance to ~2^-(8G-2) of the PRIOR variance, not of the variance itself: with G = 6 that is ~1e-12 of the prior
    # (5e-11 of the variance over the search box, 2e-9 ... 1.2e-8 exactly AT the C1-C3 training points, where it has collapsed
    # by four orders of magnitude); every extra slice buys 2^-8.  `_select_int8` therefore measures, per fitted model, the
    # ABSOLUTE accuracy delta_G of the int8 variances (as a fraction of the prior) against the FP64 contraction on a probe set
    # built around the training points (the worst cancellation) plus uniform box points, and derives the variance-collapse
    # limit `ratio_limit = delta_G / INT8_PROBE_TOL`: a point whose variance is at least that fraction of the prior is within
    # 2.5e-10 relatively (4x inside the north-star bar of 1e-9); points below it are recognised at run time (status-word byte
    # of `sample_reduce`, or the returned variances of `model.posterior`) and re-evaluated through the FP64 contraction.  The
    # SMALLEST forward slice count whose limit leaves all uniform box probes on the int8 path is used (C1-C3: 6 -- ordinary
    # points keep > 2 % of the prior variance, the limit is ~1 %), then the smallest backward count that keeps the variance
    # gradients (singles and near pairs) within 1e-7 of their largest component; a model no ladder entry serves runs on 'dmma'.
    G_FWD_LADDER = (6, 7)
    G_BWD_LADDER = (5, 6, 7)
    INT8_PROBE_TOL = 2.5e-10
    INT8_PROBE_GRAD_TOL = 1e-7
    # q-batches whose conditioning byte floor(-4 log2 rho) exceeds this go through the FP64 contraction (see
    # acquisition/_fused.py::fused_acquisition): 16 <=> rho < 1/16, i.e. some point of the q-batch keeps less than 6 % of its
    # posterior variance once the baseline draws and the preceding points are known
    INT8_COND_LIMIT = 16

    def _probe_points(self, Xt: Tensor) -> Tensor:
        """Training points (smallest posterior variances, hence the worst cancellation), the same points displaced by
        1e-6 and 1e-3 of the bounding box, and uniform points of the box -- in the raw input space, P x 1 x d."""
        n = Xt.shape[0]
        take = torch.linspace(0, n - 1, min(n, 192), device=Xt.device).round().long()
        g = torch.Generator(device="cpu").manual_seed(0)
        lo, hi = Xt.min(dim=0).values, Xt.max(dim=0).values
        width = (hi - lo).clamp_min(1e-12)
        T = Xt[take]
        dirs = torch.randn(T.shape[0], self.d, generator=g, dtype=torch.float64).to(Xt.device)
        dirs = dirs / dirs.norm(dim=-1, keepdim=True)
        box = lo + width * torch.rand(256, self.d, generator=g, dtype=torch.float64).to(Xt.device)
        pts = torch.cat([T, T + 1e-6 * width * dirs, T + 1e-3 * width * dirs, box])
        return (pts * self.x_coef + self.x_offset).unsqueeze(1)

    def _set_slices(self, g_fwd: int, g_bwd: int) -> None:
        if self.Rt_slices is None or self.Rt_slices.shape[0] != g_fwd:
            self.Rt_slices, self.Rt_scale = self._slice_rows(self.Rt, g_fwd)
        if self.R_slices is None or self.R_slices.shape[0] != g_bwd:
            self.R_slices, self.R_scale = self._slice_rows(self.R, g_bwd)
        self.g_fwd, self.g_bwd = g_fwd, g_bwd
        self.desc.g_fwd, self.desc.g_bwd = g_fwd, g_bwd
        self.desc.Rt_slices, self.desc.Rt_scale = self.Rt_slices.data_ptr(), self.Rt_scale.data_ptr()
        self.desc.R_slices, self.desc.R_scale = self.R_slices.data_ptr(), self.R_scale.data_ptr()
        self.desc.contraction = 1

    def _int8_exact(self, g: int) -> bool:
        # every diagonal of slice products must stay exact in int32 (ozaki_imma.cu: K * 128^2 * G < 2^31)
        return self.np * 16384 * g < 2**31

    def _select_int8(self, Xt: Tensor) -> None:
        """Pick (g_fwd, g_bwd) and the variance-collapse limit for this model, or fall back to 'dmma' (see the comment above)."""
        import math

        forced = settings.int8_slices.value()
        if forced is not None:  # developer / benchmark override: fixed slice counts, no probe, no variance-collapse re-routing
            if not (self._int8_exact(max(forced))):
                raise _lib.McacqError("settings.int8_slices: train set too large for exact int32 slice products")
            self._set_slices(int(forced[0]), int(forced[1]))
            return
        probe = self._probe_points(Xt)
        n_box = 256
        self.desc.contraction = 0
        v64, g64 = self._probe_variance_and_grad(probe)
        prior = self.y_std * self.y_std * self.outputscale
        ratio64 = (v64 / prior).clamp_min(0.0)          # how far the variance has collapsed at each probe point
        gmax = g64.abs().max().clamp_min(1e-300)
        chosen_f = None
        err = gerr = float("inf")
        for gf in self.G_FWD_LADDER:
            if not self._int8_exact(gf):
                break
            self._set_slices(gf, self.G_BWD_LADDER[0])
            v8, _ = self._probe_variance_and_grad(probe, backward=False)
            delta = float(((v8 - v64).abs() / prior).max())          # absolute accuracy as a fraction of the prior variance
            # every point whose variance is at least `ratio_limit` of the prior is then within INT8_PROBE_TOL relatively (which
            # is itself 4x inside the 1e-9 bar); points below it are re-evaluated through the FP64 contraction at run time
            ratio_limit = min(1.0, delta / self.INT8_PROBE_TOL)
            keep = ratio64 >= ratio_limit
            rel = (v8 - v64).abs() / v64.abs().clamp_min(1e-300)
            # usable only if the ordinary evaluation points (the uniform box probes) stay on the int8 path
            box_ok = bool(keep[-n_box:].all())
            err = float(rel[keep].max()) if (box_ok and bool(keep.any())) else float(rel[-n_box:].max())
            if os.environ.get("MCACQ_PROBE_DEBUG"):
                print(f"[int8 probe] G_fwd={gf}: delta={delta:.3e} ratio_limit={ratio_limit:.3e} box ratio min={float(ratio64[-n_box:].min()):.3e} "
                      f"train ratio min={float(ratio64[:-n_box].min()):.3e} box_ok={box_ok} err={err:.3e} "
                      f"box rel max={float(rel[-n_box:].max()):.3e} argmax delta idx={int(((v8 - v64).abs()).argmax())} of {probe.shape[0]}")
            if box_ok and err <= self.INT8_PROBE_TOL:
                chosen_f = gf
                self.int8_var_ratio_limit = ratio_limit
                # status-word byte floor(-4 log2 ratio): ratio < limit  =>  byte >= floor(-4 log2 limit)
                self.int8_var_byte_limit = (int(math.floor(-4.0 * math.log2(ratio_limit))) - 1) if ratio_limit > 0.0 else 255
                break
        chosen_b = None
        if chosen_f is not None:
            keep_rows = (ratio64 >= self.int8_var_ratio_limit)
            pairs, pair_gc = self._pair_probe(probe)
            self.desc.contraction = 0
            _, gp64 = self._probe_blocks(pairs, pair_gc)
            gpmax = gp64.abs().max().clamp_min(1e-300)
            for gb in self.G_BWD_LADDER:
                if not self._int8_exact(gb):
                    break
                self._set_slices(chosen_f, gb)
                _, g8 = self._probe_variance_and_grad(probe)
                _, gp8 = self._probe_blocks(pairs, pair_gc)
                gerr = max(float((g8 - g64)[keep_rows].abs().max() / gmax) if bool(keep_rows.any()) else 0.0,
                           float((gp8 - gp64).abs().max() / gpmax))
                if gerr <= self.INT8_PROBE_GRAD_TOL:
                    chosen_b = gb
                    break
        self.int8_probe_error, self.int8_probe_grad_error = err, gerr
        if chosen_f is not None and chosen_b is not None:
            self._set_slices(chosen_f, chosen_b)
            return
        why = ("train set too large for exact int32 slice products" if not self._int8_exact(self.G_FWD_LADDER[0]) else
               f"on the probe set the posterior variance differs by {err:.1e} and its gradient by {gerr:.1e} (relative) from "
               "the FP64 contraction (ill-conditioned train covariance)")
        warnings.warn(f"int8 contraction disabled for this model: {why}; using 'dmma'.", NumericalWarning, stacklevel=3)
        self.contraction = "dmma"
        self.int8_var_ratio_limit, self.int8_var_byte_limit = None, None
        self.Rt_slices = self.Rt_scale = self.R_slices = self.R_scale = None
        self.desc.contraction = 0
        self.desc.Rt_slices = self.desc.Rt_scale = self.desc.R_slices = self.desc.R_scale = None

    def _probe_variance_and_grad(self, probe: Tensor, backward: bool = True) -> tuple[Tensor, Tensor | None]:
        """Posterior variance at P single points and d(sum of variances)/dX through the forward AND backward contraction of
        the current mode (the backward one uses fewer slices, so it is the first to lose digits)."""
        P = probe.shape[0]
        covar, gX = self._probe_blocks(probe, torch.ones(P, 1, 1, device=self.device, dtype=torch.float64) if backward else None)
        return covar.reshape(-1), (gX.reshape(P, self.d) if backward else None)

    def _probe_blocks(self, X: Tensor, gcov: Tensor | None) -> tuple[Tensor, Tensor | None]:
        """Posterior covariance blocks of P q-batches and, with a cotangent `gcov` [P x q x q], d<gcov, covar>/dX."""
        P, q = X.shape[0], X.shape[1]
        L, st = _lib.lib(), _lib.stream_ptr()
        f64 = dict(device=self.device, dtype=torch.float64)
        X = X.contiguous()
        mean, covar, gX = torch.empty(P, q, **f64), torch.empty(P, q, q, **f64), torch.empty(P, q, self.d, **f64)
        ws = self.workspace(P, q, 0)
        _lib.check(L.mcacq_posterior(C.byref(self.desc), X.data_ptr(), P, q, mean.data_ptr(), covar.data_ptr(), ws.data_ptr(),
                                     ws.numel(), st), "mcacq_posterior (probe)")
        if gcov is None:
            return covar, None
        gm, gc = torch.zeros(P, q, **f64), gcov.contiguous()
        _lib.check(L.mcacq_posterior_backward(C.byref(self.desc), X.data_ptr(), P, q, gm.data_ptr(), gc.data_ptr(),
                                              gX.data_ptr(), ws.data_ptr(), ws.numel(), st), "mcacq_posterior_backward (probe)")
        return covar, gX

    def _pair_probe(self, probe: Tensor) -> tuple[Tensor, Tensor]:
        """The cotangent a nearly singular q x q conditional covariance sends backwards: pairs of points 1e-2 of the box apart
        with gcov = [[1, -1], [-1, 1]], i.e. the gradient of Var(f(x) - f(x')) -- a Schur-complement pivot, whose terms
        cancel almost completely.  Returns (pairs P x 2 x d, gcov P x 2 x 2)."""
        n_t = (probe.shape[0] - 256) // 3
        a, bpt = probe[:n_t, 0], probe[2 * n_t:3 * n_t, 0]        # training points and their 1e-3 displacements
        pairs = torch.stack([a, a + 10.0 * (bpt - a)], dim=1)
        gc = torch.tensor([[1.0, -1.0], [-1.0, 1.0]], device=self.device, dtype=torch.float64).expand(pairs.shape[0], 2, 2)
        return pairs.contiguous(), gc.contiguous()

    def _slice_rows(self, Mx: Tensor, G: int) -> tuple[Tensor, Tensor]:
        rows, K = Mx.shape
        S = torch.empty(G, rows, K, dtype=torch.int8, device=self.device)
        scale = torch.empty(rows, dtype=torch.float64, device=self.device)
        _lib.check(_lib.lib().mcacq_slice_rows(Mx.data_ptr(), rows, K, K, K, G, 0, 0, S.data_ptr(), scale.data_ptr(),
                                               _lib.stream_ptr()), "slice_rows")
        return S, scale

    # ---------------------------------------------------------------- helpers over the C ABI
    def workspace(self, b: int, q: int, r: int = 0) -> Tensor:
        nbytes = _lib.lib().mcacq_workspace_bytes_model(C.byref(self.desc), b, q, r)
        if nbytes == 0 and b > 0:
            raise _lib.McacqError("mcacq_workspace_bytes_model: invalid model descriptor or shape")
        return torch.empty(nbytes, dtype=torch.uint8, device=self.device)

    def scale(self, X: Tensor) -> Tensor:
        X = X.to(device=self.device, dtype=torch.float64).contiguous()
        U = torch.empty_like(X)
        _lib.check(_lib.lib().mcacq_scale_inputs(X.data_ptr(), X.numel() // self.d, self.d, self.x_offset.data_ptr(),
                                                 self.x_coef.data_ptr(), self.lengthscale.data_ptr(), U.data_ptr(),
                                                 _lib.stream_ptr()), "scale_inputs")
        return U

    def contracted_rows(self, U: Tensor) -> Tensor:
        """A = K(U, U_train) R for a (rows x d) block of scaled inputs -> rows x np."""
        L = _lib.lib()
        rows = U.shape[0]
        f64 = dict(device=self.device, dtype=torch.float64)
        Kt = torch.empty(rows, self.np, **f64)
        A = torch.empty(rows, self.np, **f64)
        counter = torch.zeros(64, dtype=torch.int32, device=self.device)
        st = _lib.stream_ptr()
        _lib.check(L.mcacq_cov_cross(self.kernel_id, self.outputscale, U.data_ptr(), rows, self.U_train.data_ptr(),
                                     self.n, self.d, Kt.data_ptr(), self.np, st), "cov_cross")
        _lib.check(L.mcacq_dgemm_tri(_lib.TRI_UPPER, rows, self.np, Kt.data_ptr(), self.R.data_ptr(), A.data_ptr(),
                                     counter.data_ptr(), st), "dgemm_tri")
        return A

    def posterior_blocks(self, X: Tensor) -> tuple[Tensor, Tensor]:
        """X: b x q x d -> (mean b x q, covar b x q x q) on the original outcome scale."""
        b, q, d = X.shape
        X = X.to(device=self.device, dtype=torch.float64).contiguous()
        f64 = dict(device=self.device, dtype=torch.float64)
        mean = torch.empty(b, q, **f64)
        covar = torch.empty(b, q, q, **f64)
        ws = self.workspace(b, q, 0)
        _lib.check(_lib.lib().mcacq_posterior(C.byref(self.desc), X.data_ptr(), b, q, mean.data_ptr(), covar.data_ptr(),
                                              ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "mcacq_posterior")
        return mean, covar

    def joint_posterior(self, X: Tensor) -> tuple[Tensor, Tensor]:
        """Joint posterior over N points (N may exceed the fused kernels' q limit): mean [N], covar [N x N].

        Setup-time route (qLogNEI baseline, prune_inferior_points, Thompson-style candidate sets): the
        cross-covariance and the contraction `A = K(X, X_train) R` run in the hand-written kernels; the
        N x N Gram `A A^T` is one library matmul."""
        X = X.to(device=self.device, dtype=torch.float64).reshape(-1, self.d).contiguous()
        N = X.shape[0]
        L = _lib.lib()
        st = _lib.stream_ptr()
        U = self.scale(X)
        f64 = dict(device=self.device, dtype=torch.float64)
        A = torch.empty(N, self.np, **f64)
        counter = torch.zeros(64, dtype=torch.int32, device=self.device)
        if self.contraction == "int8":
            # the contraction on the INT8 tensor cores with the most accurate slice counts (FP64-level: 2^-54 of the prior): the
            # covariance kernel emits the slices and the partial sums of K alpha directly (as in the fused forward pass)
            v = self.max_slices_view()
            G = v.g_fwd
            _, e = math.frexp(self.outputscale)
            n_tiles = (self.np + 63) // 64
            n_parts = (n_tiles + 7) // 8
            slices = torch.empty(G, N, self.np, dtype=torch.int8, device=self.device)
            mean_part = torch.empty(n_tiles * N, **f64)
            _lib.check(L.mcacq_cov_cross_sliced(self.kernel_id, self.outputscale, U.data_ptr(), N, self.U_train.data_ptr(), self.n,
                                                self.d, self.np, self.alpha.data_ptr(), G, e, slices.data_ptr(),
                                                mean_part.data_ptr(), st), "cov_cross_sliced")
            scale = torch.full((N,), 2.0 ** (e + 2), **f64)
            _lib.check(L.mcacq_ozaki_contract(_lib.TRI_UPPER, N, self.np, self.np, G, slices.data_ptr(), scale.data_ptr(),
                                              v.Rt_slices.data_ptr(), v.Rt_scale.data_ptr(), A.data_ptr(), self.np, st),
                       "ozaki_contract")
            k_alpha = mean_part[: n_parts * N].view(n_parts, N).sum(dim=0)
        else:
            Kt = torch.empty(N, self.np, **f64)
            _lib.check(L.mcacq_cov_cross(self.kernel_id, self.outputscale, U.data_ptr(), N, self.U_train.data_ptr(),
                                         self.n, self.d, Kt.data_ptr(), self.np, st), "cov_cross")
            _lib.check(L.mcacq_dgemm_tri(_lib.TRI_UPPER, N, self.np, Kt.data_ptr(), self.R.data_ptr(), A.data_ptr(),
                                         counter.data_ptr(), st), "dgemm_tri")
            k_alpha = Kt @ self.alpha
        Kxx = torch.empty(N, N, **f64)
        _lib.check(L.mcacq_cov_cross(self.kernel_id, self.outputscale, U.data_ptr(), N, U.data_ptr(), N, self.d,
                                     Kxx.data_ptr(), N, st), "cov_cross")
        if self.contraction == "int8" and N >= 1024 and settings.int8_gram.value():
            # large candidate sets: the Gram A A^T on the INT8 tensor cores as well (slices of A against themselves, dense mode;
            # the full square at ~100 TF/s fp64-equivalent beats the symmetric half on the 35 TF/s DMMA pipe)
            As, a_scale = self._slice_rows(A, G)
            Gm = torch.empty(N, N, **f64)
            _lib.check(L.mcacq_ozaki_contract(_lib.TRI_DENSE, N, N, self.np, G, As.data_ptr(), a_scale.data_ptr(), As.data_ptr(),
                                              a_scale.data_ptr(), Gm.data_ptr(), N, st), "ozaki_contract (gram)")
            del As
            Kxx.sub_(Gm).mul_(self.y_std**2)   # (the int8 Gram is EXACTLY symmetric: integer slice products, commutative scales)
        else:
            # covar = s^2 (K(X, X) - A A^T) in place: lower tiles only, mirrored stores (csrc/dgemm_nt.cu, mode 2)
            _lib.check(L.mcacq_syrk_sub(N, self.np, A.data_ptr(), self.np, Kxx.data_ptr(), N, self.y_std**2, counter.data_ptr(), st),
                       "syrk_sub")
        mean = self.y_mean + self.y_std * (self.mean_const + k_alpha)
        return mean, Kxx

    def batched_joint_posterior(self, X: Tensor, max_rows: int = 1 << 16) -> tuple[Tensor, Tensor]:
        """Differentiable joint posteriors of B point sets at once: X [B x N x d] -> mean [B x N], covar [B x N x N]
        (`_BatchedJointPosterior`), chunked so that one pass holds at most ~`max_rows` rows of the contraction."""
        B, N, d = X.shape
        step = max(1, min(max_rows // max(N, 1), (1 << 27) // max(N * N * d, 1)))   # rows of A, and the N x N x d differences
        if B <= step:
            return _BatchedJointPosterior.apply(X, self)
        ms, cs = zip(*(_BatchedJointPosterior.apply(X[i:i + step], self) for i in range(0, B, step)))
        return torch.cat(ms), torch.cat(cs)

    def joint_posterior_with_grad(self, X: Tensor) -> tuple[Tensor, Tensor]:
        """Differentiable `joint_posterior` (N x d -> mean [N], covar [N x N]) for joint posteriors beyond the fused kernels'
        q limit, e.g. the generic qLogNEI route over cat[X, X_baseline] when r > MCACQ_MAX_R."""
        return _JointPosterior.apply(X, self)

    def lower_times_samples(self, chol: Tensor, Z: Tensor) -> Tensor:
        """Y[N x S] = L[N x N] Z[S x N]^T with the triangular-aware DMMA kernel (MultivariateNormal.rsample's
        `root @ base_samples`)."""
        N, S = chol.shape[-1], Z.shape[0]
        chol = chol.contiguous()
        Z = Z.contiguous()
        Y = torch.empty(N, S, device=self.device, dtype=torch.float64)
        if S <= 8:
            # a handful of samples: memory-bound sweep over the triangle (one warp per row), any N
            _lib.check(_lib.lib().mcacq_lower_times_few(N, S, chol.data_ptr(), N, Z.data_ptr(), N, Y.data_ptr(), S,
                                                        _lib.stream_ptr()), "lower_times_few")
            return Y
        if N % 2:
            raise _lib.McacqError("lower_times_samples needs an even number of points (16-byte row chunks)")
        counter = torch.zeros(64, dtype=torch.int32, device=self.device)
        _lib.check(_lib.lib().mcacq_dgemm_nt(1, N, S, N, chol.data_ptr(), N, Z.data_ptr(), N, Y.data_ptr(), S,
                                             counter.data_ptr(), _lib.stream_ptr()), "dgemm_nt (trmm)")
        return Y


def _self_kernel_blocks(U3: Tensor, kernel_id: int, outputscale: float) -> Tensor:
    """k(u_i, u_j) inside every t-batch of scaled points U3 [B x N x d] -> [B x N x N], from direct differences (like the CUDA
    covariance kernels: no GEMM-expansion cancellation); differentiable torch ops, Matern gradient zero at coincident points
    (the reference's `clamp_min(1e-30).sqrt()`)."""
    diff = U3.unsqueeze(-2) - U3.unsqueeze(-3)
    sq = diff.square().sum(dim=-1)
    if kernel_id == 0:
        return outputscale * torch.exp(-0.5 * sq)
    r = sq.clamp_min(1e-30).sqrt()
    s5r = math.sqrt(5.0) * r
    return outputscale * (1.0 + s5r + (5.0 / 3.0) * sq) * torch.exp(-s5r)


class _BatchedJointPosterior(torch.autograd.Function):
    """Joint posteriors of B point sets of N points each in ONE pass: X [B x N x d] -> mean [B x N], covar [B x N x N].

    The route of everything the fused kernels do not cover (N = q + r > 32 points per t-batch: generic qLogNEI with a custom
    objective or a baseline beyond 64 points, `model.posterior` on large q): the cross-covariance against the training set
    and the contraction `A = K R` run ONCE over all B N rows in the hand-written kernels, the per-set Grams `A_b A_b^T` are one
    batched library matmul, the small within-set kernel blocks are torch ops.  Backward: dA = -s^2 (G + G^T) A per set (bmm),
    dKt = dA R^T (DMMA kernel), dU from `mcacq_cov_cross_bwd` with the rank-1 mean term, plus autograd through the
    within-set blocks.  (Round 1 looped over the t-batch in Python: ~8 launches per q-batch.)"""

    @staticmethod
    def forward(ctx, X: Tensor, strat: "DevicePredictionStrategy"):
        B, N, d = X.shape
        Xc = X.detach().to(device=strat.device, dtype=torch.float64).reshape(-1, d).contiguous()
        M = Xc.shape[0]
        L, st = _lib.lib(), _lib.stream_ptr()
        f64 = dict(device=strat.device, dtype=torch.float64)
        U = strat.scale(Xc)
        Kt = torch.empty(M, strat.np, **f64)
        _lib.check(L.mcacq_cov_cross(strat.kernel_id, strat.outputscale, U.data_ptr(), M, strat.U_train.data_ptr(), strat.n,
                                     strat.d, Kt.data_ptr(), strat.np, st), "cov_cross")
        A = torch.empty(M, strat.np, **f64)
        counter = torch.zeros(64, dtype=torch.int32, device=strat.device)
        _lib.check(L.mcacq_dgemm_tri(_lib.TRI_UPPER, M, strat.np, Kt.data_ptr(), strat.R.data_ptr(), A.data_ptr(),
                                     counter.data_ptr(), st), "dgemm_tri")
        mean = (strat.y_mean + strat.y_std * (strat.mean_const + Kt @ strat.alpha)).view(B, N)
        del Kt
        A3 = A.view(B, N, strat.np)
        Kxx = _self_kernel_blocks(U.view(B, N, d), strat.kernel_id, strat.outputscale)
        covar = (strat.y_std**2) * (Kxx - torch.bmm(A3, A3.mT))
        ctx.strat = strat
        ctx.save_for_backward(U, A)
        ctx.dims = (B, N, d)
        return mean, covar

    @staticmethod
    def backward(ctx, gmean: Tensor, gcovar: Tensor):
        U, A = ctx.saved_tensors
        strat = ctx.strat
        B, N, d = ctx.dims
        M = B * N
        L, st = _lib.lib(), _lib.stream_ptr()
        f64 = dict(device=strat.device, dtype=torch.float64)
        gm = torch.zeros(M, **f64) if gmean is None else gmean.to(**f64).reshape(M).contiguous()
        gc = torch.zeros(B, N, N, **f64) if gcovar is None else gcovar.to(**f64)
        s2 = strat.y_std**2
        dA = (-s2 * torch.bmm(gc + gc.mT, A.view(B, N, strat.np))).reshape(M, strat.np).contiguous()
        dKt = torch.empty(M, strat.np, **f64)
        counter = torch.zeros(64, dtype=torch.int32, device=strat.device)
        _lib.check(L.mcacq_dgemm_tri(_lib.TRI_LOWER, M, strat.np, dA.data_ptr(), strat.Rt.data_ptr(), dKt.data_ptr(),
                                     counter.data_ptr(), st), "dgemm_tri")
        del dA
        rs = (strat.y_std * gm).contiguous()
        dU = torch.empty(M, strat.d, **f64)
        _lib.check(L.mcacq_cov_cross_bwd(strat.kernel_id, strat.outputscale, U.data_ptr(), M, strat.U_train.data_ptr(), strat.n,
                                         strat.d, dKt.data_ptr(), strat.np, rs.data_ptr(), strat.alpha.data_ptr(),
                                         dU.data_ptr(), 0, st), "cov_cross_bwd(train)")
        del dKt
        with torch.enable_grad():
            U3 = U.view(B, N, d).detach().requires_grad_(True)
            Kxx = _self_kernel_blocks(U3, strat.kernel_id, strat.outputscale)
            (gU,) = torch.autograd.grad(Kxx, U3, grad_outputs=s2 * gc)
        dU = dU + gU.reshape(M, d)
        return (dU / (strat.x_coef * strat.lengthscale)).view(B, N, d), None


class _JointPosterior(torch.autograd.Function):
    """Joint posterior over N points with a hand-assembled backward from the same C entry points as the fused path:
    dA = -s^2 (G + G^T) A,  dKt = dA R^T (+ s gmean alpha^T inside the covariance backward),  dU from `mcacq_cov_cross_bwd`
    against the train inputs and against the points themselves (the K(X, X) term).  Setup-time / fallback route: the two
    N x N products are library matmuls."""

    @staticmethod
    def forward(ctx, X: Tensor, strat: "DevicePredictionStrategy"):
        Xc = X.detach().to(device=strat.device, dtype=torch.float64).reshape(-1, strat.d).contiguous()
        N = Xc.shape[0]
        L, st = _lib.lib(), _lib.stream_ptr()
        f64 = dict(device=strat.device, dtype=torch.float64)
        U = strat.scale(Xc)
        Kt = torch.empty(N, strat.np, **f64)
        _lib.check(L.mcacq_cov_cross(strat.kernel_id, strat.outputscale, U.data_ptr(), N, strat.U_train.data_ptr(), strat.n,
                                     strat.d, Kt.data_ptr(), strat.np, st), "cov_cross")
        A = torch.empty(N, strat.np, **f64)
        counter = torch.zeros(64, dtype=torch.int32, device=strat.device)
        _lib.check(L.mcacq_dgemm_tri(_lib.TRI_UPPER, N, strat.np, Kt.data_ptr(), strat.R.data_ptr(), A.data_ptr(),
                                     counter.data_ptr(), st), "dgemm_tri")
        Kxx = torch.empty(N, N, **f64)
        _lib.check(L.mcacq_cov_cross(strat.kernel_id, strat.outputscale, U.data_ptr(), N, U.data_ptr(), N, strat.d,
                                     Kxx.data_ptr(), N, st), "cov_cross")
        mean = strat.y_mean + strat.y_std * (strat.mean_const + Kt @ strat.alpha)
        covar = (strat.y_std**2) * (Kxx - A @ A.mT)
        ctx.strat = strat
        ctx.save_for_backward(U, A)
        return mean, covar

    @staticmethod
    def backward(ctx, gmean: Tensor, gcovar: Tensor):
        U, A = ctx.saved_tensors
        strat = ctx.strat
        N = U.shape[0]
        L, st = _lib.lib(), _lib.stream_ptr()
        f64 = dict(device=strat.device, dtype=torch.float64)
        gm = torch.zeros(N, **f64) if gmean is None else gmean.to(**f64).contiguous()
        gc = torch.zeros(N, N, **f64) if gcovar is None else gcovar.to(**f64)
        s2 = strat.y_std**2
        Gs = (s2 * (gc + gc.mT)).contiguous()
        dA = -(Gs @ A).contiguous()
        dKt = torch.empty(N, strat.np, **f64)
        counter = torch.zeros(64, dtype=torch.int32, device=strat.device)
        _lib.check(L.mcacq_dgemm_tri(_lib.TRI_LOWER, N, strat.np, dA.data_ptr(), strat.Rt.data_ptr(), dKt.data_ptr(),
                                     counter.data_ptr(), st), "dgemm_tri")
        rs = (strat.y_std * gm).contiguous()
        dU = torch.empty(N, strat.d, **f64)
        _lib.check(L.mcacq_cov_cross_bwd(strat.kernel_id, strat.outputscale, U.data_ptr(), N, strat.U_train.data_ptr(), strat.n,
                                         strat.d, dKt.data_ptr(), strat.np, rs.data_ptr(), strat.alpha.data_ptr(),
                                         dU.data_ptr(), 0, st), "cov_cross_bwd(train)")
        _lib.check(L.mcacq_cov_cross_bwd(strat.kernel_id, strat.outputscale, U.data_ptr(), N, U.data_ptr(), N, strat.d,
                                         Gs.data_ptr(), N, None, None, dU.data_ptr(), 1, st), "cov_cross_bwd(self)")
        return dU / (strat.x_coef * strat.lengthscale), None
