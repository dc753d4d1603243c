"""qLogExpectedImprovement / qLogNoisyExpectedImprovement behind the reference's interface
(botorch/acquisition/logei.py:72-585, 688-725).

`forward(X: (b) x q x d) -> (b)` runs as ONE fused CUDA pipeline (cross-covariance -> FP64 DMMA
contraction against the cached factor -> posterior blocks -> Cholesky + reparameterised samples +
log_fatplus / fatmax / logmeanexp) with a hand-written backward, whenever the configuration is the
default one: exact `SingleTaskGP`, identity objective, no posterior transform, no outcome constraints,
a normal base-sample sampler, and (for qLogNEI) the cached-root path.  Anything else takes the generic
torch-op route of `SampleReducingMCAcquisitionFunction`, which still gets its posterior from the CUDA
kernels.
"""
from __future__ import annotations

import warnings
from copy import deepcopy
from functools import partial
from typing import Callable

import torch
from torch import Tensor

from .. import _lib
from ..exceptions.errors import BotorchError, NanError, NotPSDError
from ..exceptions.warnings import BotorchWarning
from ..models.gp_regression import SingleTaskGP
from ..sampling.base import MCSampler
from ..sampling.get_sampler import get_sampler
from ..sampling.normal import NormalMCSampler
from ..utils.safe_math import fatmax, log_fatplus, log_softplus, logmeanexp, smooth_amax
from ..utils.transforms import concatenate_pending_points, match_batch_shape, t_batch_mode_transform
from ._fused import BaselineOperands, MCOperands, affine_constraints, fused_acquisition
from .cached_cholesky import CachedCholeskyMCSamplerMixin
from .monte_carlo import SampleReducingMCAcquisitionFunction
from .objective import LinearMCObjective, MCAcquisitionObjective, PosteriorTransform
from .utils import compute_best_feasible_objective, prune_inferior_points

TAU_RELU = 1e-6
TAU_MAX = 1e-2


class _ShapePosterior:
    """Shape/device stand-in handed to samplers so they build base samples without a posterior call."""

    def __init__(self, batch_shape: torch.Size, q: int, device, dtype) -> None:
        self.batch_shape = torch.Size(batch_shape)
        self.base_sample_shape = self.batch_shape + torch.Size([q])
        self.batch_range = (0, -1)
        self.device, self.dtype = device, dtype


class LogImprovementMCAcquisitionFunction(SampleReducingMCAcquisitionFunction):
    _log: bool = True

    def __init__(self, model, sampler: MCSampler | None = None, objective: MCAcquisitionObjective | None = None,
                 posterior_transform: PosteriorTransform | None = None, X_pending: Tensor | None = None,
                 constraints: list[Callable[[Tensor], Tensor]] | None = None, eta: Tensor | float = 1e-3,
                 fat: bool = True, tau_max: float = TAU_MAX) -> None:
        q_reduction = partial(fatmax if fat else smooth_amax, tau=tau_max)
        super().__init__(model=model, sampler=sampler, objective=objective, posterior_transform=posterior_transform,
                         X_pending=X_pending, sample_reduction=logmeanexp, q_reduction=q_reduction,
                         constraints=constraints, eta=eta, fat=fat)
        self.tau_max = tau_max
        self._mc_cache: dict = {}
        self._utility_mode: int | None = None  # set by the non-log subclasses (qEI / qNEI / qSR / qPI)

    # ---- fused-route plumbing -------------------------------------------------------------------
    def _fusable(self, X: Tensor) -> bool:
        return (isinstance(self.model, SingleTaskGP) and self._affine_objective() is not None
                and self.posterior_transform is None and self._fused_constraints() is not None
                and X.is_cuda and X.dtype in (torch.float64, torch.float32)
                and X.shape[-2] <= _lib.MAX_Q and len(self.sample_shape) == 1
                and (self.sampler is None or isinstance(self.sampler, NormalMCSampler)))

    def _affine_objective(self) -> tuple[float, float] | None:
        """(weight, offset) when the objective is affine in the single outcome: `IdentityMCObjective` or a
        `LinearMCObjective` with one weight (reference objective.py:314-358); None otherwise."""
        if self._identity_objective:
            return 1.0, 0.0
        if isinstance(self.objective, LinearMCObjective) and self.objective.weights.numel() == 1:
            return float(self.objective.weights.reshape(-1)[0]), 0.0
        return None

    def _fused_constraints(self) -> tuple | None:
        if not hasattr(self, "_fused_cons"):
            self._fused_cons = affine_constraints(self._constraints, self._eta, self._fat,
                                                  device=self.model.train_inputs[0].device)
        return self._fused_cons

    def _mc_extras(self) -> dict:
        w, o = self._affine_objective()
        return dict(obj_weight=w, obj_offset=o, constraints=self._fused_constraints(), con_fat=bool(self._fat),
                    mode=self._utility_mode, util_param=float(getattr(self, "_util_param", 0.0)))

    def _ensure_sampler(self, probe) -> None:
        if self.sampler is None:
            self.sampler = get_sampler(posterior=probe, sample_shape=self._default_sample_shape)


class qLogExpectedImprovement(LogImprovementMCAcquisitionFunction):
    """MC-based batch Log Expected Improvement (reference :143-243)."""

    def __init__(self, model, best_f: float | Tensor, sampler: MCSampler | None = None,
                 objective: MCAcquisitionObjective | None = None, posterior_transform: PosteriorTransform | None = None,
                 X_pending: Tensor | None = None, constraints=None, eta: Tensor | float = 1e-3, fat: bool = True,
                 tau_max: float = TAU_MAX, tau_relu: float = TAU_RELU) -> None:
        super().__init__(model=model, sampler=sampler, objective=objective, posterior_transform=posterior_transform,
                         X_pending=X_pending, constraints=constraints, eta=eta, tau_max=check_tau(tau_max, name="tau_max"),
                         fat=fat)
        self.register_buffer("best_f", torch.as_tensor(best_f))
        self.tau_relu = check_tau(tau_relu, name="tau_relu")

    def _sample_forward(self, obj: Tensor) -> Tensor:
        return _log_improvement(Y=obj, best_f=self.best_f, tau=self.tau_relu, fat=self._fat)

    def _mc_operands(self, X: Tensor) -> MCOperands:
        q = X.shape[-2]
        key = (q, X.device)
        ops = self._mc_cache.get(key)
        if ops is None:
            probe = _ShapePosterior(X.shape[:-2], q, X.device, X.dtype)
            self._ensure_sampler(probe)
            self.sampler._construct_base_samples(posterior=probe)
            S = self.sample_shape.numel()
            Zt = self.sampler.base_samples.reshape(S, q).t().contiguous()
            if self.best_f.numel() != 1:
                raise BotorchError("The fused qLogEI route expects a scalar `best_f`.")
            best = torch.full((S,), float(self.best_f), device=X.device, dtype=torch.float64)
            ops = MCOperands(Zt=Zt, best=best, tau_relu=float(self.tau_relu), tau_max=float(self.tau_max), fat=self._fat,
                             **self._mc_extras())
            self._mc_cache = {key: ops}
        return ops

    @concatenate_pending_points
    @t_batch_mode_transform()
    def forward(self, X: Tensor) -> Tensor:
        if not self._fusable(X) or self.best_f.numel() != 1:
            return self._sample_reduction(self._q_reduction(self._non_reduced_forward(X=X)))
        strat = self.model.prediction_strategy()
        batch_shape = X.shape[:-2]
        Xf = X.reshape(-1, *X.shape[-2:]).to(torch.float64)  # float32 callers: computed in fp64, returned in their dtype
        out = _chunked_fused(Xf, strat, None, self._mc_operands(Xf))
        return out.reshape(batch_shape).to(X.dtype)


class qLogNoisyExpectedImprovement(LogImprovementMCAcquisitionFunction, CachedCholeskyMCSamplerMixin):
    """MC-based batch Log Noisy Expected Improvement (reference :246-585)."""

    def __init__(self, model, X_baseline: Tensor, sampler: MCSampler | None = None,
                 objective: MCAcquisitionObjective | None = None, posterior_transform: PosteriorTransform | None = None,
                 X_pending: Tensor | None = None, constraints=None, eta: Tensor | float = 1e-3, fat: bool = True,
                 prune_baseline: bool = True, cache_root: bool | None = None, tau_max: float = TAU_MAX,
                 tau_relu: float = TAU_RELU, marginalize_dim: int | None = None, incremental: bool = True) -> None:
        self.incremental = incremental
        super().__init__(model=model, sampler=sampler, objective=objective, posterior_transform=posterior_transform,
                         X_pending=None if incremental else X_pending, constraints=constraints, eta=eta, fat=fat,
                         tau_max=check_tau(tau_max, name="tau_max"))
        if incremental:
            self.X_pending = None  # the attribute optimize_acqf / concatenate_pending_points read (reference :364)
        self.tau_relu = check_tau(tau_relu, name="tau_relu")
        self.prune_baseline = prune_baseline
        self.marginalize_dim = marginalize_dim
        self._init_baseline(model=model, X_baseline=X_baseline, X_pending=X_pending, sampler=sampler,
                            objective=objective, posterior_transform=posterior_transform, cache_root=cache_root)

    # ---- baseline -------------------------------------------------------------------------------
    def _init_baseline(self, model, X_baseline: Tensor, X_pending: Tensor | None = None, sampler=None, objective=None,
                       posterior_transform=None, cache_root: bool | None = None) -> None:
        CachedCholeskyMCSamplerMixin.__init__(self, model=model, cache_root=cache_root, sampler=sampler)
        if self.prune_baseline:
            X_baseline = prune_inferior_points(model=model, X=X_baseline, objective=objective,
                                               posterior_transform=posterior_transform,
                                               marginalize_dim=self.marginalize_dim, constraints=self._constraints)
        self.register_buffer("_X_baseline", X_baseline)
        if X_pending is not None and self.incremental:
            full_X_baseline = torch.cat([X_baseline, X_pending], dim=-2)
        else:
            full_X_baseline = X_baseline
        self.register_buffer("_full_X_baseline", full_X_baseline)
        self.register_buffer("baseline_samples", None)
        self.register_buffer("baseline_obj", None)
        self._mc_cache = {}
        self._base_ops = None
        if self._cache_root:
            self.q_in = -1
            with torch.no_grad():
                posterior = self.model.posterior(self.X_baseline, posterior_transform=self.posterior_transform)
                self.baseline_samples = self.get_posterior_samples(posterior)
                self.baseline_obj = self.objective(self.baseline_samples, X=self.X_baseline)
            self.base_sampler = deepcopy(self.sampler)
            self.register_buffer("_baseline_best_f", self._compute_best_feasible_objective(
                samples=self.baseline_samples, obj=self.baseline_obj))
            self._baseline_L = self._compute_root_decomposition(posterior=posterior)

    @property
    def X_baseline(self) -> Tensor:
        return self._full_X_baseline

    def set_X_pending(self, X_pending: Tensor | None = None) -> None:
        if not self.incremental:
            return super().set_X_pending(X_pending=X_pending)
        if X_pending is None:
            if not hasattr(self, "_full_X_baseline") or self._full_X_baseline.shape[-2] == self._X_baseline.shape[-2]:
                return
        self._init_baseline(model=self.model, X_baseline=self._X_baseline, X_pending=X_pending, sampler=self.sampler,
                            objective=self.objective, posterior_transform=self.posterior_transform,
                            cache_root=self._cache_root)

    def compute_best_f(self, obj: Tensor) -> Tensor:
        if self._cache_root:
            val = self._baseline_best_f
        else:
            val = self._compute_best_feasible_objective(samples=self.baseline_samples, obj=self.baseline_obj)
        n_sample_dims = len(self.sample_shape)
        view_shape = torch.Size([*val.shape[:n_sample_dims], *(1,) * (obj.ndim - val.ndim - 1), *val.shape[n_sample_dims:]])
        return val.view(view_shape).to(obj)

    def _compute_best_feasible_objective(self, samples: Tensor, obj: Tensor) -> Tensor:
        return compute_best_feasible_objective(samples=samples, obj=obj, constraints=self._constraints, model=self.model,
                                               objective=self.objective, posterior_transform=self.posterior_transform,
                                               X_baseline=self.X_baseline)

    def _sample_forward(self, obj: Tensor) -> Tensor:
        return _log_improvement(Y=obj, best_f=self.compute_best_f(obj), tau=self.tau_relu, fat=self._fat)

    # ---- generic (unfused) route ----------------------------------------------------------------
    def _get_samples_and_objectives(self, X: Tensor) -> tuple[Tensor, Tensor]:
        n_baseline, q = self.X_baseline.shape[-2], X.shape[-2]
        X_full = torch.cat([match_batch_shape(self.X_baseline, X), X], dim=-2)
        posterior = self.model.posterior(X_full, posterior_transform=self.posterior_transform)
        if not self._cache_root:
            samples_full = super().get_posterior_samples(posterior)
            obj_full = self.objective(samples_full, X=X_full)
            split_dim = len(obj_full.shape) - 1
            self.baseline_samples, samples = samples_full.split([n_baseline, q], dim=split_dim)
            self.baseline_obj, obj = obj_full.split([n_baseline, q], dim=split_dim)
            return samples, obj
        self._set_sampler(q_in=q, posterior=posterior)
        samples = self._get_f_X_samples(posterior=posterior, q_in=q)
        obj = self.objective(samples, X=X_full[..., -q:, :])
        return samples, obj

    # ---- fused route ----------------------------------------------------------------------------
    def _baseline_operands(self) -> BaselineOperands:
        if self._base_ops is None:
            strat = self.model.prediction_strategy()
            Xb = self.X_baseline.to(device=strat.device, dtype=torch.float64)
            U_base = strat.scale(Xb)
            A_base = strat.contracted_rows(U_base)
            self._base_ops = BaselineOperands(U_base=U_base, A_base=A_base,
                                              L_base=self._baseline_L.to(torch.float64).contiguous())
        return self._base_ops

    def _mc_operands(self, X: Tensor) -> MCOperands:
        q, r = X.shape[-2], self.X_baseline.shape[-2]
        key = (q, X.device)
        ops = self._mc_cache.get(key)
        if ops is None:
            probe = _ShapePosterior(X.shape[:-2], r + q, X.device, X.dtype)
            self.q_in = -1
            self._set_sampler(q_in=q, posterior=probe)  # first r columns <- base_sampler (cached_cholesky.py:172-192)
            S = self.sample_shape.numel()
            Zt = self.sampler.base_samples.reshape(S, r + q).t().contiguous()
            best = self._baseline_best_f.reshape(S).to(device=X.device, dtype=torch.float64).contiguous()
            ops = MCOperands(Zt=Zt, best=best, tau_relu=float(self.tau_relu), tau_max=float(self.tau_max), fat=self._fat,
                             **self._mc_extras())
            self._mc_cache = {key: ops}
        return ops

    @concatenate_pending_points
    @t_batch_mode_transform()
    def forward(self, X: Tensor) -> Tensor:
        fused = (self._fusable(X) and self._cache_root and hasattr(self, "_baseline_L")
                 and self.X_baseline.dim() == 2
                 and _lib.fused_supported(X.shape[-2], self.X_baseline.shape[-2], self.sample_shape.numel()))
        if not fused:
            return self._sample_reduction(self._q_reduction(self._non_reduced_forward(X=X)))
        strat = self.model.prediction_strategy()
        batch_shape = X.shape[:-2]
        Xf = X.reshape(-1, *X.shape[-2:]).to(torch.float64)
        try:
            out = _chunked_fused(Xf, strat, self._baseline_operands(), self._mc_operands(Xf))
        except (NanError, NotPSDError):
            # reference: cached_cholesky.py:143-170 -- warn and fall back to joint sampling
            warnings.warn("Low-rank cholesky updates failed due NaNs or due to an ill-conditioned covariance matrix. "
                          "Falling back to standard sampling.", BotorchWarning, stacklevel=2)
            cache_root, self._cache_root = self._cache_root, False
            try:
                return self._sample_reduction(self._q_reduction(self._non_reduced_forward(X=X)))
            finally:
                self._cache_root = cache_root
        return out.reshape(batch_shape).to(X.dtype)


def _chunked_fused(Xf: Tensor, strat, base, mc, max_rows: int = 1 << 17) -> Tensor:
    """Evaluate the fused op over t-batch chunks so one workspace stays below ~2*max_rows*np doubles."""
    b, q, _ = Xf.shape
    step = max(1, max_rows // q)
    if b <= step:
        return fused_acquisition(Xf, strat, base, mc)
    return torch.cat([fused_acquisition(Xf[i:i + step], strat, base, mc) for i in range(0, b, step)])


def _log_improvement(Y: Tensor, best_f: Tensor, tau: float | Tensor, fat: bool) -> Tensor:
    """log of the softplus-smoothed improvement (reference :688-715)."""
    log_soft_clamp = log_fatplus if fat else log_softplus
    Z = Y - best_f.unsqueeze(-1).to(Y)
    return log_soft_clamp(Z, tau=tau)


def check_tau(tau, name: str):
    """Validity check of the temperature arguments (reference :718-725)."""
    if isinstance(tau, Tensor) and tau.numel() != 1:
        raise ValueError(f"{name} is not a scalar: {tau.numel()=}.")
    if not (tau > 0):
        raise ValueError(f"{name} is non-positive: {tau=}.")
    return tau
