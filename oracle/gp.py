"""Exact-GP posterior restatement (gpytorch ExactGP eval path as configured by BoTorch).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED against gpytorch itself (pinned to
scikit-learn's GaussianProcessRegressor instead, tests/test_oracle_vs_sklearn.py): gpytorch>=1.15.2 and
linear_operator>=0.6.1 are not vendored under /root/reference and not installable here; the
functions below restate their published algorithms as BoTorch configures them:

* BoTorch global settings  botorch/__init__.py:52-57  (fast computations off,
  cholesky_max_tries=6, max_cholesky_size=4096, max_eager_kernel_size=4096).
* Per-posterior settings   botorch/models/utils/assorted.py:305-315  (fast_pred_var on,
  detach_test_caches on unless propagate_grads).
* Call sites               botorch/models/gpytorch.py:544-610 (posterior),
  botorch/models/gp_regression.py:209-214 (forward: mean_module + covar_module),
  botorch/models/transforms/input.py:541-554 (Normalize), outcome.py:431-511 (Standardize).

gpytorch pieces restated (names refer to gpytorch/linear_operator modules):
  kernels.kernel.sq_dist / dist, RBFKernel.forward, MaternKernel.forward, ScaleKernel.forward,
  models.exact_prediction_strategies.DefaultPredictionStrategy.{mean_cache, covar_cache,
  exact_predictive_mean, exact_predictive_covar}, linear_operator.utils.cholesky.psd_safe_cholesky,
  LinearOperator.root_inv_decomposition(method="cholesky"), distributions.MultivariateNormal.rsample.
"""
from __future__ import annotations

import math
import warnings
from dataclasses import dataclass, field

import torch
from torch import Tensor


class NanError(RuntimeError):
    """linear_operator.utils.errors.NanError stand-in."""


class NotPSDError(RuntimeError):
    """linear_operator.utils.errors.NotPSDError stand-in."""


class NumericalWarning(RuntimeWarning):
    """linear_operator.utils.warnings.NumericalWarning stand-in."""


# --------------------------------------------------------------------------- kernels
def sq_dist(x1: Tensor, x2: Tensor, x1_eq_x2: bool = False) -> Tensor:
    """gpytorch.kernels.kernel.sq_dist: GEMM expansion after centring on x1.mean(-2)."""
    adjustment = x1.mean(-2, keepdim=True)
    x1 = x1 - adjustment
    x1_norm = x1.pow(2).sum(dim=-1, keepdim=True)
    x1_pad = torch.ones_like(x1_norm)
    if x1_eq_x2 and not x1.requires_grad and not x2.requires_grad:
        x2, x2_norm, x2_pad = x1, x1_norm, x1_pad
    else:
        x2 = x2 - adjustment
        x2_norm = x2.pow(2).sum(dim=-1, keepdim=True)
        x2_pad = torch.ones_like(x2_norm)
    x1_ = torch.cat([-2.0 * x1, x1_norm, x1_pad], dim=-1)
    x2_ = torch.cat([x2, x2_pad, x2_norm], dim=-1)
    res = x1_.matmul(x2_.transpose(-2, -1))
    if x1_eq_x2 and not x1.requires_grad and not x2.requires_grad:
        res.diagonal(dim1=-2, dim2=-1).fill_(0)
    return res.clamp_min(0)


def dist(x1: Tensor, x2: Tensor, x1_eq_x2: bool = False) -> Tensor:
    """gpytorch.kernels.kernel.dist: sqrt(clamp_min(sq_dist, 1e-30))."""
    return sq_dist(x1, x2, x1_eq_x2=x1_eq_x2).clamp_min(1e-30).sqrt()


def rbf_forward(x1: Tensor, x2: Tensor, lengthscale: Tensor) -> Tensor:
    """gpytorch RBFKernel.forward (ARD branch): exp(-0.5 * sq_dist(x1/l, x2/l))."""
    x1_eq_x2 = torch.equal(x1, x2)
    x1_ = x1.div(lengthscale)
    x2_ = x2.div(lengthscale)
    return sq_dist(x1_, x2_, x1_eq_x2=x1_eq_x2).div(-2).exp()


def matern52_forward(x1: Tensor, x2: Tensor, lengthscale: Tensor) -> Tensor:
    """gpytorch MaternKernel.forward, nu=2.5: mean of all x1 rows subtracted first."""
    x1_eq_x2 = torch.equal(x1, x2)
    mean = x1.reshape(-1, x1.size(-1)).mean(0)[(None,) * (x1.dim() - 1)]
    x1_ = (x1 - mean).div(lengthscale)
    x2_ = (x2 - mean).div(lengthscale)
    distance = dist(x1_, x2_, x1_eq_x2=x1_eq_x2)
    exp_component = torch.exp(-math.sqrt(5.0) * distance)
    constant_component = (math.sqrt(5) * distance).add(1).add(5.0 / 3.0 * distance**2)
    return constant_component * exp_component


def psd_safe_cholesky(A: Tensor, max_tries: int = 6, jitter: float | None = None) -> Tensor:
    """linear_operator.utils.cholesky.psd_safe_cholesky: jitter 1e-8*10^i (fp64) / 1e-6*10^i
    (fp32) added to FAILING batch elements only, up to `max_tries` times."""
    L, info = torch.linalg.cholesky_ex(A)
    if not torch.any(info):
        return L
    if torch.isnan(A).any():
        raise NanError(f"cholesky_cpu: {int(torch.isnan(A).sum())} of {A.numel()} elements are NaN.")
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


# --------------------------------------------------------------------------- model
@dataclass
class OracleGP:
    """Fitted-state container + exact prediction, single output (m=1), no model batch.

    train_X is the RAW training input (d columns); Normalize (offset, coef) and Standardize
    (Y mean m, stdv s) follow the reference's transform semantics.  kernel in {"rbf","matern52"};
    outputscale None == no ScaleKernel (SingleTaskGP default, gp_regression.py:192-196).
    """

    train_X: Tensor  # n x d (raw)
    train_Y: Tensor  # n x 1 (raw)
    lengthscale: Tensor  # d
    noise: Tensor  # scalar (inferred, homoskedastic) or n (fixed, already standardized units)
    kernel: str = "rbf"
    outputscale: float | None = None
    mean_constant: float = 0.0
    norm_offset: Tensor | None = None  # d   (Normalize: X' = (X - offset) / coef)
    norm_coef: Tensor | None = None  # d
    standardize: bool = True
    _cache: dict = field(default_factory=dict, repr=False)

    # -- transforms
    def transform_inputs(self, X: Tensor) -> Tensor:
        """models/model.py:197-217 + transforms/input.py:541-554."""
        if self.norm_offset is None:
            return X
        return (X - self.norm_offset) / self.norm_coef

    def _y_stats(self):
        """transforms/outcome.py:340-352: nanstd (unbiased) with min_stdv=1e-8; nanmean."""
        if not self.standardize:
            return torch.zeros((), dtype=self.train_Y.dtype), torch.ones((), dtype=self.train_Y.dtype)
        Y = self.train_Y
        if Y.shape[-2] == 1:
            stdv = torch.ones((), dtype=Y.dtype)
        else:
            # models/transforms/utils.py:146-160 (nanstd): sqrt(mean((Y-mean)^2) * n / (n-1))
            nobs = Y.shape[-2]
            stdv = ((Y - Y.mean(dim=-2, keepdim=True)).pow(2).mean(dim=-2) * nobs / (nobs - 1)).sqrt().squeeze()
        stdv = torch.where(stdv >= 1e-8, stdv, torch.ones_like(stdv))
        return Y.mean(dim=-2).squeeze(), stdv

    # -- kernel
    def k(self, x1: Tensor, x2: Tensor) -> Tensor:
        ls = self.lengthscale.view(*([1] * (x1.dim() - 1)), -1)
        if self.kernel == "rbf":
            out = rbf_forward(x1, x2, ls)
        elif self.kernel == "matern52":
            out = matern52_forward(x1, x2, ls)
        else:
            raise ValueError(self.kernel)
        if self.outputscale is not None:
            out = out.mul(self.outputscale)  # ScaleKernel.forward
        return out

    # -- caches (gpytorch DefaultPredictionStrategy)
    def caches(self):
        if "L" not in self._cache:
            with torch.no_grad():
                Xt = self.transform_inputs(self.train_X)
                m, s = self._y_stats()
                y = ((self.train_Y - m) / s).squeeze(-1)
                K = self.k(Xt, Xt)
                noise = self.noise if self.noise.dim() == 0 else self.noise.reshape(-1)
                Khat = K + torch.diag_embed(noise.expand(K.shape[-1]).to(K))
                L = psd_safe_cholesky(Khat)
                # mean_cache = (K + s2 I)^{-1} (y - mu)  via cholesky_solve
                mean_cache = torch.cholesky_solve((y - self.mean_constant).unsqueeze(-1), L).squeeze(-1)
                # covar_cache = root_inv_decomposition().root = (L^{-1})^T  (fast_pred_var on)
                eye = torch.eye(L.shape[-1], dtype=L.dtype)
                Linv = torch.linalg.solve_triangular(L, eye, upper=False)
                self._cache.update(Xt=Xt, L=L, mean_cache=mean_cache, covar_cache=Linv.mT.contiguous(), m=m, s=s)
        c = self._cache
        return c["Xt"], c["L"], c["mean_cache"], c["covar_cache"], c["m"], c["s"]

    # -- posterior
    def posterior_mvn(self, X: Tensor):
        """Returns (mean [..., q], covar [..., q, q]) on the ORIGINAL outcome scale.

        gpytorch ExactGP.__call__ (eval) -> exact_prediction; eager joint-kernel branch when
        n + q <= max_eager_kernel_size (4096), else two lazy slices.  Then
        Standardize.untransform_posterior (outcome.py:479-511): mean m + s*mu, covar diag(s) S diag(s).
        """
        Xt, L, mean_cache, covar_cache, m, s = self.caches()
        Xq = self.transform_inputs(X)
        n, q = Xt.shape[-2], Xq.shape[-2]
        batch = Xq.shape[:-2]
        Xtr_b = Xt.expand(*batch, *Xt.shape)
        if n + q <= 4096:
            full = torch.cat([Xtr_b, Xq], dim=-2)
            rows = self.k(Xq, full)
            test_train, test_test = rows[..., :n], rows[..., n:]
        else:
            test_train = self.k(Xq, Xtr_b)
            test_test = self.k(Xq, Xq)
        mean = (test_train @ mean_cache.unsqueeze(-1)).squeeze(-1) + self.mean_constant
        A = test_train.matmul(covar_cache)
        covar = torch.add(test_test, A @ A.transpose(-1, -2), alpha=-1)
        mean_tf = m + s * mean
        sf = s.expand(covar.shape[:-1])
        covar_tf = sf.unsqueeze(-1) * covar * sf.unsqueeze(-2)
        return mean_tf, covar_tf


def mvn_rsample_from_base_samples(mean: Tensor, covar: Tensor, base_samples: Tensor, sample_shape) -> Tensor:
    """posteriors/gpytorch.py:86-127 + gpytorch MultivariateNormal.rsample(base_samples=...):
    root = psd_safe_cholesky(covar) (cholesky_max_tries=6); samples = root @ z + loc; output has the
    trailing output dim of size 1."""
    root = psd_safe_cholesky(covar, max_tries=6)
    bs = base_samples.reshape(-1, *mean.shape[:-1], root.shape[-1])
    bs = bs.permute(*range(1, mean.dim() + 1), 0)
    res = root.matmul(bs) + mean.unsqueeze(-1)
    res = res.permute(-1, *range(mean.dim())).contiguous()
    res = res.view(torch.Size(sample_shape) + mean.shape)
    return res.unsqueeze(-1)
