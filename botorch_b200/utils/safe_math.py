"""Numerically safe reductions for the UNFUSED fallback route of the acquisition classes.

The fused CUDA kernel (`csrc/sample_reduce.cu`) implements `log_fatplus -> fatmax -> logmeanexp`
in-kernel.  Acquisition configurations the kernel does not cover (custom objectives, constraints)
materialise samples with the CUDA posterior kernels and reduce them with the torch-op versions below,
which follow botorch/utils/safe_math.py (`_inf_max_helper` :146-191, `logmeanexp` :213-225,
`log_softplus` :228-249, `smooth_amax` :252-278, `fatplus` :307-325, `fatmax` :328-355).
"""
from __future__ import annotations

import math

import torch
from torch import Tensor
from torch.nn.functional import softplus

TAU = 1.0
ALPHA = 2.0


def _dims(dim) -> tuple[int, ...]:
    return (dim,) if isinstance(dim, int) else tuple(dim)


def _inf_max_helper(max_fun, x: Tensor, dim, keepdim: bool) -> Tensor:
    dims = _dims(dim)
    if any(x.shape[d] == 0 for d in dims):
        return x.sum(dim=dim, keepdim=keepdim) - torch.inf
    M = x.amax(dim=dim, keepdim=True)
    inf_max = M.isinf() & (x == M)
    any_inf = inf_max
    for d in dims:
        any_inf = any_inf.any(dim=d, keepdim=True)
    M_fin = M.masked_fill(M.isinf(), 0.0)
    finite_part = M_fin + max_fun(x.masked_fill(any_inf, 0.0) - M_fin, dim=dim, keepdim=True)
    inf_part = x.masked_fill(~inf_max, 0.0).sum(dim=dim, keepdim=True)
    res = torch.where(any_inf, inf_part, finite_part)
    return res if keepdim else res.sum(dim=dim)


def logsumexp(x: Tensor, dim, keepdim: bool = False) -> Tensor:
    return _inf_max_helper(torch.logsumexp, x=x, dim=dim, keepdim=keepdim)


def logmeanexp(X: Tensor, dim, keepdim: bool = False) -> Tensor:
    n = math.prod(X.shape[i] for i in _dims(dim))
    return logsumexp(X, dim=dim, keepdim=keepdim) - math.log(n)


def cauchy(x: Tensor) -> Tensor:
    return 1 / (1 + x.square())


def fatplus(x: Tensor, tau=TAU) -> Tensor:
    u = x / tau
    return tau * (softplus(u) + 1e-1 * cauchy(u))


def log_fatplus(x: Tensor, tau=TAU) -> Tensor:
    return fatplus(x, tau=tau).log()


def log_softplus(x: Tensor, tau=TAU) -> Tensor:
    tau = torch.as_tensor(tau, dtype=x.dtype, device=x.device)
    upper, lower = (16, -15) if x.dtype == torch.float32 else (32, -35)
    mask = x / tau > lower
    soft = softplus(x.masked_fill(~mask, lower), beta=(1 / tau), threshold=upper).log()
    return torch.where(mask, soft, x / tau + tau.log())


def _pareto(x: Tensor, alpha: float, check: bool = True) -> Tensor:
    if check and (x < 0).any():
        raise ValueError("Argument `x` must be non-negative.")
    a = alpha / 2
    b1 = 2 * a
    b0 = a * b1
    return (b0 / (b0 + b1 * x + x.square())).pow(a)


def fatmax(x: Tensor, dim, keepdim: bool = False, tau=TAU, alpha: float = ALPHA) -> Tensor:
    def max_fun(y: Tensor, dim, keepdim: bool = False) -> Tensor:
        return tau * _pareto(-y / tau, alpha=alpha).sum(dim=dim, keepdim=keepdim).log()

    return _inf_max_helper(max_fun=max_fun, x=x, dim=dim, keepdim=keepdim)


def smooth_amax(X: Tensor, dim=-1, keepdim: bool = False, tau=1.0) -> Tensor:
    return logsumexp(X / tau, dim=dim, keepdim=keepdim) * tau


def log1mexp(x: Tensor) -> Tensor:
    """log(1 - exp(x)) for x < 0 (reference :36-46)."""
    is_small = -math.log(2) < x
    return torch.where(is_small, (-x.expm1()).log(), (-x.exp()).log1p())


def logplusexp(a: Tensor, b: Tensor) -> Tensor:
    """log(exp(a) + exp(b)) (reference :100-103)."""
    return logsumexp(torch.stack(torch.broadcast_tensors(a, b), dim=-1), dim=-1)


def logdiffexp(log_a: Tensor, log_b: Tensor) -> Tensor:
    """log(b - a) given log a < log b (reference :106-120)."""
    log_a, log_b = torch.broadcast_tensors(log_a, log_b)
    is_inf = log_b == -torch.inf
    return log_b + log1mexp(log_a - log_b.masked_fill(is_inf, 0.0))


def log1pexp(x: Tensor) -> Tensor:
    """log(1 + exp(x)) without overflow, switching form at x = 18 (reference :84-93)."""
    small = x <= 18
    lo = x.masked_fill(~small, 0).exp().log1p()
    hi_arg = x.masked_fill(small, 0)
    return torch.where(small, lo, hi_arg + (-hi_arg).exp())


def logexpit(X: Tensor) -> Tensor:
    """log sigmoid(X) (reference :96-98)."""
    return -log1pexp(-X)


_INV_SQRT_3 = math.sqrt(1 / 3)


def fatmoid(X: Tensor, tau=1.0) -> Tensor:
    """Fat-tailed smooth step, O(1/x^2) as x -> -inf, inflection at 1/sqrt(3) (reference :441-458)."""
    u = X / tau
    return torch.where(u < 0, 2 / 3 * cauchy(u - _INV_SQRT_3), 1 - 2 / 3 * cauchy(u + _INV_SQRT_3))


def log_fatmoid(X: Tensor, tau=1.0) -> Tensor:
    return fatmoid(X, tau=tau).log()


def sigmoid(X: Tensor, log: bool = False, fat: bool = False) -> Tensor:
    """(log-)sigmoid with an optional fat tail (reference :493-507)."""
    Y = log_fatmoid(X) if fat else logexpit(X)
    return Y if log else Y.exp()
