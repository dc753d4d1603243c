"""qLogEI / qLogNEI forward restated on top of oracle.gp (TEST INFRASTRUCTURE ONLY).

Follows the reference:
  acquisition/monte_carlo.py:109-125, 268-305   sample -> objective -> _sample_forward -> reductions
  acquisition/logei.py:127-135                   q_reduction=fatmax(tau_max) / sample_reduction=logmeanexp
  acquisition/logei.py:228-243, 688-715          qLogEI._sample_forward / _log_improvement
  acquisition/logei.py:393-459                   qLogNEI._init_baseline
  acquisition/logei.py:527-565                   qLogNEI._get_samples_and_objectives
  acquisition/cached_cholesky.py:98-192          root decomposition cache, fallback, _set_sampler
  utils/low_rank.py:84-172                       sample_cached_cholesky
  acquisition/utils.py:134-138                   best_f = obj.amax(-1) under no_grad
Gradients come from autograd, exactly like generation/gen.py:466-469.
"""
from __future__ import annotations

import warnings

import torch
from torch import Tensor

from . import safe_math as sm
from .gp import NanError, NotPSDError, OracleGP, mvn_rsample_from_base_samples, psd_safe_cholesky
from .sampling import qlogei_base_samples, qlognei_base_samples


def _reduce(obj: Tensor, best_f: Tensor, tau_relu: float, tau_max: float, fat: bool) -> Tensor:
    li = sm.log_improvement(obj, best_f, tau_relu, fat)
    qred = sm.fatmax(li, dim=-1, tau=tau_max) if fat else sm.smooth_amax(li, dim=-1, tau=tau_max)
    return sm.logmeanexp(qred, dim=0)


class OracleQLogEI:
    def __init__(self, gp: OracleGP, best_f, S: int, seed: int, tau_relu=sm.TAU_RELU, tau_max=sm.TAU_MAX, fat=True):
        self.gp, self.S, self.seed = gp, S, seed
        self.best_f = torch.as_tensor(best_f)  # logei.py:226 -- a python float becomes fp32, as in the reference
        self.tau_relu, self.tau_max, self.fat = tau_relu, tau_max, fat
        self._base = {}

    def samples(self, X: Tensor) -> Tensor:
        q = X.shape[-2]
        if q not in self._base:
            self._base[q] = qlogei_base_samples(self.S, q, self.seed, X.dtype)
        mean, covar = self.gp.posterior_mvn(X)
        bs = self._base[q].expand(self.S, *X.shape[:-2], q, 1)
        return mvn_rsample_from_base_samples(mean, covar, bs, (self.S,))  # S x b x q x 1

    def __call__(self, X: Tensor) -> Tensor:
        if X.dim() == 2:
            X = X.unsqueeze(0)
        obj = self.samples(X).squeeze(-1)
        return _reduce(obj, self.best_f, self.tau_relu, self.tau_max, self.fat)


class OracleQLogNEI:
    """cache_root=True, prune_baseline=False (explicit X_baseline), IdentityMCObjective."""

    def __init__(self, gp: OracleGP, X_baseline: Tensor, S: int, seed: int, tau_relu=sm.TAU_RELU,
                 tau_max=sm.TAU_MAX, fat=True):
        self.gp, self.S, self.seed = gp, S, seed
        self.X_baseline = X_baseline
        self.tau_relu, self.tau_max, self.fat = tau_relu, tau_max, fat
        r = X_baseline.shape[-2]
        with torch.no_grad():
            mean_b, covar_b = gp.posterior_mvn(X_baseline)
            zb = qlogei_base_samples(S, r, seed, X_baseline.dtype)
            # the MVN root decomposition is cached by linear_operator and reused for _baseline_L
            self.baseline_L = psd_safe_cholesky(covar_b, max_tries=6)
            bs = zb.reshape(-1, r).t()
            base_samples = (self.baseline_L @ bs + mean_b.unsqueeze(-1)).t()  # S x r
            self.baseline_best_f = base_samples.amax(dim=-1)  # [S]
        self._base = {}
        self.fell_back = False

    def _f_X_samples(self, X: Tensor) -> Tensor:
        r, q = self.X_baseline.shape[-2], X.shape[-2]
        if q not in self._base:
            self._base[q] = qlognei_base_samples(self.S, r, q, self.seed, X.dtype)
        Z = self._base[q]  # S x 1 x (r+q) x 1
        X_full = torch.cat([self.X_baseline.expand(*X.shape[:-2], r, X.shape[-1]), X], dim=-2)
        mean, covar = self.gp.posterior_mvn(X_full)
        try:
            bottom_rows = covar[..., -q:, :]
            bl, br = bottom_rows.split([r, q], dim=-1)
            bl_chol = torch.linalg.solve_triangular(self.baseline_L, bl.transpose(-2, -1), upper=False).transpose(-2, -1)
            br_to_chol = br - bl_chol @ bl_chol.transpose(-2, -1)
            br_chol = psd_safe_cholesky(br_to_chol, max_tries=6)
            new_Lq = torch.cat([bl_chol, br_chol], dim=-1).unsqueeze(-3)  # b x 1 x q x (r+q)
            # _reshape_base_samples (low_rank.py:42-81), single output: b x 1 x (r+q) x S
            bsmp = Z.view(self.S, r + q).t().expand(*X.shape[:-2], 1, r + q, self.S)
            new_mean = mean.unsqueeze(-1)[..., -q:, :]
            res = (
                new_Lq.matmul(bsmp)
                .add(new_mean.transpose(-1, -2).unsqueeze(-1))
                .permute(-1, *range(mean.dim() - 1), -2, -3)
                .contiguous()
            )
            if torch.isnan(res).any() or torch.isinf(res).any():
                raise NanError("Samples contain nans/infs.")
            return res  # S x b x q x 1
        except (NanError, NotPSDError):
            warnings.warn("Low-rank cholesky updates failed. Falling back to standard sampling.", RuntimeWarning)
            self.fell_back = True
        bs = Z.expand(self.S, *X.shape[:-2], r + q, 1)
        samples = mvn_rsample_from_base_samples(mean, covar, bs, (self.S,))
        return samples[..., -q:, :]

    def __call__(self, X: Tensor) -> Tensor:
        if X.dim() == 2:
            X = X.unsqueeze(0)
        obj = self._f_X_samples(X).squeeze(-1)  # S x b x q
        best_f = self.baseline_best_f.view(self.S, *([1] * (obj.dim() - 2)))
        return _reduce(obj, best_f, self.tau_relu, self.tau_max, self.fat)


def value_and_grad(acqf, X: Tensor):
    """generation/gen.py:435-470: losses = f(X); grad of losses.sum() w.r.t. X."""
    Xg = X.detach().clone().requires_grad_(True)
    vals = acqf(Xg)
    (g,) = torch.autograd.grad(vals.sum(), Xg)
    return vals.detach(), g


# ---- non-log sample-reducing utilities ------------------------------------------------------
def oracle_qei(orc: OracleQLogEI, X: Tensor) -> Tensor:
    """qExpectedImprovement (monte_carlo.py:427-437 + default amax / mean reductions :268-290)."""
    obj = orc.samples(X if X.dim() > 2 else X.unsqueeze(0)).squeeze(-1)
    return (obj - orc.best_f.to(obj)).clamp_min(0).amax(dim=-1).mean(dim=0)


def oracle_qsr(orc: OracleQLogEI, X: Tensor) -> Tensor:
    """qSimpleRegret (monte_carlo.py:821-830)."""
    obj = orc.samples(X if X.dim() > 2 else X.unsqueeze(0)).squeeze(-1)
    return obj.amax(dim=-1).mean(dim=0)


def oracle_qpi(orc: OracleQLogEI, X: Tensor, tau: float = 1e-3) -> Tensor:
    """qProbabilityOfImprovement (monte_carlo.py:752-763)."""
    obj = orc.samples(X if X.dim() > 2 else X.unsqueeze(0)).squeeze(-1)
    return torch.sigmoid((obj - orc.best_f.to(obj)) / tau).amax(dim=-1).mean(dim=0)


def oracle_qnei(orc: OracleQLogNEI, X: Tensor) -> Tensor:
    """qNoisyExpectedImprovement, cached-root path (monte_carlo.py:607-616)."""
    X = X if X.dim() > 2 else X.unsqueeze(0)
    obj = orc._f_X_samples(X).squeeze(-1)
    best = orc.baseline_best_f.view(orc.S, *([1] * (obj.dim() - 1)))
    return (obj - best).clamp_min(0).amax(dim=-1).mean(dim=0)


def oracle_qucb(orc: OracleQLogEI, X: Tensor, beta: float, lower: bool = False) -> Tensor:
    """qUpperConfidenceBound / qLowerConfidenceBound (monte_carlo.py:833-921): mean + (-)sqrt(beta pi / 2) |obj - mean|."""
    import math

    obj = orc.samples(X if X.dim() > 2 else X.unsqueeze(0)).squeeze(-1)
    bp = math.sqrt(beta * math.pi / 2) * (-1.0 if lower else 1.0)
    mean = obj.mean(dim=0)
    return (mean + bp * (obj - mean).abs()).amax(dim=-1).mean(dim=0)


def oracle_qpstd(orc: OracleQLogEI, X: Tensor) -> Tensor:
    """qPosteriorStandardDeviation (monte_carlo.py:924-989)."""
    import math

    obj = orc.samples(X if X.dim() > 2 else X.unsqueeze(0)).squeeze(-1)
    return ((obj - obj.mean(dim=0)).abs() * math.sqrt(math.pi / 2)).amax(dim=-1).mean(dim=0)
