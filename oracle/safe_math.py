"""Restatement of the reduction arithmetic of the hot path.

Follows /root/reference/botorch/utils/safe_math.py (v0.18.1):
  _inf_max_helper :146-191, logsumexp :123-143, logmeanexp :213-225,
  log_softplus :228-249, smooth_amax :252-278, fatplus/log_fatplus :298-325,
  fatmax :328-355, cauchy :461-463, _pareto :466-490.
Pinned against the reference module itself in tests/test_oracle_vs_reference.py and
against tests/golden/safe_math_*.pt.
"""
from __future__ import annotations

import math

import torch
from torch import Tensor
from torch.nn.functional import softplus

TAU_RELU = 1e-6  # acquisition/logei.py:67
TAU_MAX = 1e-2  # acquisition/logei.py:68
ALPHA = 2.0  # safe_math.py:31


def _any(x: Tensor, dim, keepdim: bool = False) -> Tensor:
    dims = (dim,) if isinstance(dim, int) else tuple(dim)
    for d in dims:
        x = x.any(dim=d, keepdim=True)
    if keepdim:
        return x
    for d in sorted((d % x.ndim for d in dims), reverse=True):
        x = x.squeeze(d)
    return x


def _inf_max_helper(max_fun, x: Tensor, dim, keepdim: bool) -> Tensor:
    """safe_math.py:146-191 -- inf-aware smooth maximum skeleton."""
    dims = (dim,) if isinstance(dim, int) else tuple(dim)
    if any(x.shape[d] == 0 for d in dims):
        return x.sum(dim=dim, keepdim=keepdim) - torch.inf
    M = x.amax(dim=dim, keepdim=True)
    is_inf_max = torch.logical_and(*torch.broadcast_tensors(M.isinf(), x == M))
    has_inf_max = _any(is_inf_max, dim=dim, keepdim=True)
    y_inf = x.masked_fill(~is_inf_max, 0.0)
    M_no_inf = M.masked_fill(M.isinf(), 0.0)
    y_no_inf = x.masked_fill(has_inf_max, 0.0) - M_no_inf
    res = torch.where(
        has_inf_max,
        y_inf.sum(dim=dim, keepdim=True),
        M_no_inf + max_fun(y_no_inf, dim=dim, keepdim=True),
    )
    return res if keepdim else res.sum(dim=dim)


def logsumexp(x: Tensor, dim, keepdim: bool = False) -> Tensor:
    """safe_math.py:123-143."""
    return _inf_max_helper(torch.logsumexp, x=x, dim=dim, keepdim=keepdim)


def logmeanexp(X: Tensor, dim, keepdim: bool = False) -> Tensor:
    """safe_math.py:213-225."""
    n = X.shape[dim] if isinstance(dim, int) else math.prod(X.shape[i] for i in dim)
    return logsumexp(X, dim=dim, keepdim=keepdim) - math.log(n)


def cauchy(x: Tensor) -> Tensor:
    """safe_math.py:461-463."""
    return 1 / (1 + x.square())


def fatplus(x: Tensor, tau=1.0) -> Tensor:
    """safe_math.py:307-325: tau * (softplus(x/tau) + 0.1 * cauchy(x/tau))."""
    z = x / tau
    return tau * (softplus(z) + 1e-1 * cauchy(z))


def log_fatplus(x: Tensor, tau=1.0) -> Tensor:
    """safe_math.py:298-304."""
    return fatplus(x, tau=tau).log()


def log_softplus(x: Tensor, tau=1.0) -> Tensor:
    """safe_math.py:228-249 (non-fat variant)."""
    tau = torch.as_tensor(tau, dtype=x.dtype, device=x.device)
    upper = 16 if x.dtype == torch.float32 else 32
    lower = -15 if x.dtype == torch.float32 else -35
    mask = x / tau > lower
    return torch.where(
        mask,
        softplus(x.masked_fill(~mask, lower), beta=(1 / tau), threshold=upper).log(),
        x / tau + tau.log(),
    )


def _pareto(x: Tensor, alpha: float, check: bool = True) -> Tensor:
    """safe_math.py:466-490."""
    if check and (x < 0).any():
        raise ValueError("Argument `x` must be non-negative.")
    alpha = alpha / 2
    beta_1 = 2 * alpha
    beta_0 = alpha * beta_1
    return (beta_0 / (beta_0 + beta_1 * x + x.square())).pow(alpha)


def fatmax(x: Tensor, dim, keepdim: bool = False, tau=1.0, alpha: float = ALPHA) -> Tensor:
    """safe_math.py:328-355."""

    def max_fun(x: Tensor, dim, keepdim: bool = False) -> Tensor:
        return tau * _pareto(-x / tau, alpha=alpha).sum(dim=dim, keepdim=keepdim).log()

    return _inf_max_helper(max_fun=max_fun, x=x, dim=dim, keepdim=keepdim)


def smooth_amax(X: Tensor, dim=-1, keepdim: bool = False, tau=1.0) -> Tensor:
    """safe_math.py:252-278."""
    return logsumexp(X / tau, dim=dim, keepdim=keepdim) * tau


def log_improvement(Y: Tensor, best_f: Tensor, tau, fat: bool) -> Tensor:
    """acquisition/logei.py:688-715 (`_log_improvement`)."""
    log_soft_clamp = log_fatplus if fat else log_softplus
    Z = Y - best_f.unsqueeze(-1).to(Y)
    return log_soft_clamp(Z, tau=tau)


def log1mexp(x: Tensor) -> Tensor:
    """safe_math.py:36-46."""
    is_small = -math.log(2) < x
    return torch.where(is_small, (-x.expm1()).log(), (-x.exp()).log1p())


def logplusexp(a: Tensor, b: Tensor) -> Tensor:
    """safe_math.py:100-103."""
    return logsumexp(torch.stack(torch.broadcast_tensors(a, b), dim=-1), dim=-1)


def logdiffexp(log_a: Tensor, log_b: Tensor) -> Tensor:
    """safe_math.py:106-120."""
    log_a, log_b = torch.broadcast_tensors(log_a, log_b)
    is_inf = log_b == -torch.inf
    return log_b + log1mexp(log_a - log_b.masked_fill(is_inf, 0.0))
