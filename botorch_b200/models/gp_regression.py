"""`SingleTaskGP` with a device-resident exact posterior (reference: botorch/models/gp_regression.py:53-214,
botorch/models/gpytorch.py:544-610).

Same constructor and `posterior` contract as the reference; hyper-parameters are held by the light modules
of `kernels.py` (fitting is out of scope).  The first `posterior` / acquisition call after construction or a
hyper-parameter change builds the caches on the GPU (`DevicePredictionStrategy`); every evaluation then runs
the CUDA kernels of `csrc/` through the C ABI.
"""
from __future__ import annotations

import ctypes as C
import warnings

import torch
from torch import Tensor
from torch.nn import Module

from .. import _lib, settings
from ..exceptions.errors import InputDataError, UnsupportedError
from ..exceptions.warnings import InputDataWarning
from ..posteriors.gpytorch import GPyTorchPosterior, MultivariateNormal
from .kernels import (ConstantMean, FixedNoiseGaussianLikelihood, GaussianLikelihood, Kernel, ScaleKernel,
                      get_covar_module_with_dim_scaled_prior)
from .model import Model
from .prediction_strategy import DevicePredictionStrategy
from .transforms.outcome import Standardize

DEFAULT = object()


class _PosteriorBlocks(torch.autograd.Function):
    """Differentiable (mean, covar) = posterior(X) through mcacq_posterior / mcacq_posterior_backward."""

    @staticmethod
    def forward(ctx, X: Tensor, strat: DevicePredictionStrategy):
        b, q, d = X.shape
        Xc = X.detach().contiguous()
        f64 = dict(device=X.device, dtype=torch.float64)
        mean = torch.empty(b, q, **f64)
        covar = torch.empty(b, q, q, **f64)
        ws = strat.workspace(b, q, 0)
        _lib.check(_lib.lib().mcacq_posterior(C.byref(strat.desc), Xc.data_ptr(), b, q, mean.data_ptr(),
                                              covar.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr()),
                   "mcacq_posterior")
        ctx.strat, ctx.ws = strat, ws
        ctx.save_for_backward(Xc)
        return mean, covar

    @staticmethod
    def backward(ctx, gmean: Tensor, gcovar: Tensor):
        (Xc,) = ctx.saved_tensors
        if ctx.ws is None:
            raise RuntimeError("posterior backward called twice: the workspace is consumed in place.")
        b, q, d = Xc.shape
        gm = (torch.zeros(b, q, device=Xc.device, dtype=torch.float64) if gmean is None else gmean).contiguous()
        gc = (torch.zeros(b, q, q, device=Xc.device, dtype=torch.float64) if gcovar is None else gcovar).contiguous()
        gX = torch.empty_like(Xc)
        _lib.check(_lib.lib().mcacq_posterior_backward(C.byref(ctx.strat.desc), Xc.data_ptr(), b, q, gm.data_ptr(),
                                                       gc.data_ptr(), gX.data_ptr(), ctx.ws.data_ptr(), ctx.ws.numel(),
                                                       _lib.stream_ptr()), "mcacq_posterior_backward")
        ctx.ws = None
        return gX, None


class SingleTaskGP(Model):
    def __init__(self, train_X: Tensor, train_Y: Tensor, train_Yvar: Tensor | None = None, likelihood: Module | None = None,
                 covar_module: Module | None = None, mean_module: Module | None = None, outcome_transform=DEFAULT,
                 input_transform: Module | None = None) -> None:
        super().__init__()
        if train_X.dim() != 2 or train_Y.dim() != 2:
            raise UnsupportedError("botorch_b200.SingleTaskGP supports un-batched `n x d` / `n x 1` training data.")
        if train_Y.shape[-1] != 1:
            raise UnsupportedError("botorch_b200.SingleTaskGP supports a single output (m = 1).")
        if train_X.shape[0] != train_Y.shape[0]:
            raise InputDataError("Expected X and Y to have the same number of rows.")
        if torch.isnan(train_X).any() or torch.isnan(train_Y).any():
            raise InputDataError("Input data contains NaN values.")
        if train_X.dtype != torch.float64:
            warnings.warn("The model inputs are of type {}. It is strongly recommended to use double precision in "
                          "BoTorch; botorch_b200 computes in float64.".format(train_X.dtype), InputDataWarning)
        self.train_X_raw = train_X.detach().to(torch.float64)
        self.train_Y_raw = train_Y.detach().to(torch.float64)
        d = train_X.shape[-1]
        if outcome_transform is DEFAULT:
            outcome_transform = Standardize(m=1)
        if input_transform is not None:
            input_transform.train()
            with torch.no_grad():
                transformed_X = input_transform(self.train_X_raw)
            input_transform.eval()
            self.input_transform = input_transform
        else:
            transformed_X = self.train_X_raw
        Y, Yvar = self.train_Y_raw, (None if train_Yvar is None else train_Yvar.detach().to(torch.float64))
        if outcome_transform is not None:
            outcome_transform.to(Y)
            outcome_transform.train()
            Y, Yvar = outcome_transform(Y, Yvar)
            outcome_transform.eval()
            self.outcome_transform = outcome_transform
        self.train_inputs = (transformed_X,)
        self.train_targets = Y.squeeze(-1)
        if likelihood is None:
            likelihood = GaussianLikelihood() if Yvar is None else FixedNoiseGaussianLikelihood(Yvar.squeeze(-1))
        self.likelihood = likelihood
        self.mean_module = ConstantMean() if mean_module is None else mean_module
        self.covar_module = get_covar_module_with_dim_scaled_prior(ard_num_dims=d) if covar_module is None else covar_module
        self._strategy: DevicePredictionStrategy | None = None
        self._strategy_key = None
        self.to(train_X.device)

    # ------------------------------------------------------------------ reference surface
    @property
    def batch_shape(self) -> torch.Size:
        return torch.Size()

    @property
    def num_outputs(self) -> int:
        return 1

    def train(self, mode: bool = True):
        if mode:
            self._strategy = None  # gpytorch clears the prediction strategy on .train()
        return super().train(mode)

    # ------------------------------------------------------------------ device caches
    def _hyper_key(self):
        """Identity of the fitted state WITHOUT touching device memory: storage pointer + in-place version counter of every
        hyper-parameter tensor (an optimiser round calls this twice; copying the values to the host would cost a
        device synchronisation per hyper-parameter per call)."""
        base = self._base_kernel()
        tensors = [base.lengthscale, self.likelihood.noise, self.mean_module.constant, self.train_inputs[0], self.train_targets]
        if isinstance(self.covar_module, ScaleKernel):
            tensors.append(self.covar_module.outputscale)
        ident = tuple((id(t), t.data_ptr(), t._version, tuple(t.shape)) for t in tensors)
        # the tensors are kept alive next to the cached strategy (`_strategy_tensors`), so neither their Python identity nor
        # their storage address can be recycled by a replacement tensor while the key is in use
        self._key_tensors = tensors
        return (self.train_inputs[0].device, settings.contraction.value(), settings.int8_slices.value(),
                settings.int8_cond_limit.value(), ident)

    def _base_kernel(self) -> Kernel:
        k = self.covar_module.base_kernel if isinstance(self.covar_module, ScaleKernel) else self.covar_module
        if not isinstance(k, Kernel) or k.kernel_id < 0:
            raise UnsupportedError(f"Unsupported covariance module {type(self.covar_module).__name__}: botorch_b200 "
                                   "implements ARD RBF and Matern-5/2 kernels (optionally inside a ScaleKernel).")
        return k

    def prediction_strategy(self) -> DevicePredictionStrategy:
        """Build (or return the cached) device caches: chol(K + noise), mean_cache, covar_cache = L^{-T}."""
        key = self._hyper_key()
        if self._strategy is not None and self._strategy_key == key:
            return self._strategy
        Xt = self.train_inputs[0]
        if Xt.device.type != "cuda":
            raise _lib.McacqError("SingleTaskGP must live on a CUDA device (`model.to('cuda')`); botorch_b200 has no "
                                  "CPU path.")
        d = Xt.shape[-1]
        base = self._base_kernel()
        ls = base.lengthscale.detach().reshape(-1)
        ls = ls.expand(d) if ls.numel() == 1 else ls
        tf = getattr(self, "input_transform", None)
        f64 = dict(device=Xt.device, dtype=torch.float64)
        offset = tf.offset.reshape(-1).to(**f64) if tf is not None else torch.zeros(d, **f64)
        coef = tf.coefficient.reshape(-1).to(**f64) if tf is not None else torch.ones(d, **f64)
        otf = getattr(self, "outcome_transform", None)
        y_mean = float(otf.means.reshape(-1)[0]) if otf is not None else 0.0
        y_std = float(otf.stdvs.reshape(-1)[0]) if otf is not None else 1.0
        outputscale = float(self.covar_module.outputscale.detach()) if isinstance(self.covar_module, ScaleKernel) else 1.0
        self._strategy = DevicePredictionStrategy(
            train_X_transformed=Xt, train_Y_standardized=self.train_targets, lengthscale=ls.to(**f64),
            noise=self.likelihood.noise.detach().to(**f64), kernel_id=base.kernel_id, outputscale=outputscale,
            mean_const=float(self.mean_module.constant.detach()), x_offset=offset, x_coef=coef, y_mean=y_mean, y_std=y_std,
            contraction=settings.contraction.value())
        self._strategy_key = key
        self._strategy_tensors = self._key_tensors
        return self._strategy

    def _apply(self, fn, recurse=True):
        out = super()._apply(fn, recurse)
        self.train_X_raw = fn(self.train_X_raw)
        self.train_Y_raw = fn(self.train_Y_raw)
        self.train_inputs = tuple(fn(t) for t in self.train_inputs)
        self.train_targets = fn(self.train_targets)
        self._strategy = None
        return out

    # ------------------------------------------------------------------ posterior
    def posterior(self, X: Tensor, output_indices=None, observation_noise=False, posterior_transform=None):
        """X: (batch_shape) x q x d -> GPyTorchPosterior over the q points (joint), original outcome scale."""
        if output_indices is not None:
            raise UnsupportedError("output_indices is not supported for single-output models.")
        self.eval()
        strat = self.prediction_strategy()
        if X.dim() < 2:
            raise InputDataError("X must be at least two-dimensional (q x d).")
        batch_shape, q, d = X.shape[:-2], X.shape[-2], X.shape[-1]
        Xf = X.reshape(-1, q, d).to(device=strat.device, dtype=torch.float64)
        if q > _lib.MAX_Q:
            # joint posteriors beyond the fused kernels' q limit (baseline sets, candidate sets, cat[X, X_baseline] of the generic
            # qLogNEI route): through the covariance / contraction kernels
            if Xf.shape[0] == 1 or q * q * d > (1 << 24):
                # one (or a few) LARGE point sets: per set through the DMMA SYRK-sub kernel, no N x N x d intermediates
                if X.requires_grad and torch.is_grad_enabled():
                    ms, cs = zip(*(strat.joint_posterior_with_grad(x) for x in Xf))
                else:
                    ms, cs = zip(*(strat.joint_posterior(x) for x in Xf))
                mean, covar = torch.stack(ms), torch.stack(cs)
            else:
                mean, covar = strat.batched_joint_posterior(Xf)   # all t-batches in one pass (no Python loop over b)
        else:
            mean, covar = self._posterior_chunks(Xf, strat)
        if isinstance(observation_noise, Tensor) or observation_noise:
            s2 = strat.y_std**2
            if isinstance(observation_noise, Tensor):
                # the reference adds the tensor through the likelihood BEFORE `untransform_posterior`
                # (models/gpytorch.py:517-527, 603-605), i.e. it is noise on the standardised scale
                covar = covar + s2 * torch.diag_embed(observation_noise.reshape(*covar.shape[:-1]).to(covar))
            else:
                covar = covar + torch.diag_embed((self.likelihood.noise.mean() * s2).expand(covar.shape[:-1]).to(covar))
        mean = mean.reshape(*batch_shape, q)
        covar = covar.reshape(*batch_shape, q, q)
        posterior = GPyTorchPosterior(MultivariateNormal(mean, covar))
        if posterior_transform is not None:
            return posterior_transform(posterior)
        return posterior

    def _posterior_chunks(self, Xf: Tensor, strat: DevicePredictionStrategy, max_rows: int = 1 << 17):
        """Chunk the t-batch so a workspace never exceeds ~2 * max_rows * np doubles."""
        b, q, _ = Xf.shape
        step = max(1, max_rows // q)
        if b <= step:
            return self._posterior_blocks_checked(Xf, strat)
        means, covars = [], []
        for i in range(0, b, step):
            m, c = self._posterior_blocks_checked(Xf[i:i + step], strat)
            means.append(m)
            covars.append(c)
        return torch.cat(means), torch.cat(covars)

    @staticmethod
    def _posterior_blocks_checked(Xf: Tensor, strat: DevicePredictionStrategy):
        """`_PosteriorBlocks` plus the int8 mode's variance-collapse rule (DevicePredictionStrategy._select_int8): q-batches with
        a point whose variance has dropped below `int8_var_ratio_limit` of the prior are re-evaluated through the FP64
        contraction, so `posterior.variance` keeps its relative accuracy at (and next to) training points."""
        mean, covar = _PosteriorBlocks.apply(Xf, strat)
        lim = strat.int8_var_ratio_limit if strat.contraction == "int8" else None
        if lim is None or Xf.shape[0] == 0:
            return mean, covar
        prior = strat.y_std * strat.y_std * strat.outputscale
        low = covar.detach().diagonal(dim1=-1, dim2=-2).amin(dim=-1) < lim * prior
        if bool(low.any()):
            idx = low.nonzero().squeeze(-1)
            m64, c64 = _PosteriorBlocks.apply(Xf.index_select(0, idx), strat.fp64_view())
            mean, covar = mean.index_copy(0, idx, m64), covar.index_copy(0, idx, c64)
        return mean, covar
