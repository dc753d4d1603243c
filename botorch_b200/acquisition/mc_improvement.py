"""Non-log sample-reducing MC acquisition functions on the same fused kernel:
qExpectedImprovement (botorch/acquisition/monte_carlo.py:351-437), qNoisyExpectedImprovement (:440-673),
qProbabilityOfImprovement (:676-763), qSimpleRegret (:766-830).  They differ from qLogEI / qLogNEI only in the
per-sample utility and the reductions (`torch.amax` over q, `torch.mean` over the MC samples), which are template-free
modes of `csrc/sample_reduce.cu`; everything upstream (covariance, contraction, posterior blocks, Cholesky, samples) is
shared.  Anything the fused kernel does not cover falls back to the generic torch-op route."""
from __future__ import annotations

import math
from functools import partial

import torch
from torch import Tensor

from .logei import qLogExpectedImprovement, qLogNoisyExpectedImprovement


def _use_plain_reductions(acqf) -> None:
    sample_dim = tuple(range(len(acqf.sample_shape)))
    acqf._sample_reduction = partial(torch.mean, dim=sample_dim)
    acqf._q_reduction = partial(torch.amax, dim=-1)
    acqf._fat = False  # constraint weighting of the non-log family uses the plain sigmoid (reference monte_carlo.py:146-212)


class qExpectedImprovement(qLogExpectedImprovement):
    """MC-based batch Expected Improvement: mean_S max_q relu(y - best_f)."""

    _log = False

    def __init__(self, model, best_f, sampler=None, objective=None, posterior_transform=None, X_pending=None,
                 constraints=None, eta=1e-3) -> None:
        super().__init__(model=model, best_f=best_f, sampler=sampler, objective=objective,
                         posterior_transform=posterior_transform, X_pending=X_pending, constraints=constraints, eta=eta)
        _use_plain_reductions(self)
        self._utility_mode = 2

    def _sample_forward(self, obj: Tensor) -> Tensor:
        return (obj - self.best_f.unsqueeze(-1).to(obj)).clamp_min(0)


class qSimpleRegret(qLogExpectedImprovement):
    """MC-based batch Simple Regret: mean_S max_q y."""

    _log = False

    def __init__(self, model, sampler=None, objective=None, posterior_transform=None, X_pending=None) -> None:
        super().__init__(model=model, best_f=0.0, sampler=sampler, objective=objective,
                         posterior_transform=posterior_transform, X_pending=X_pending)
        _use_plain_reductions(self)
        self._utility_mode = 3

    def _sample_forward(self, obj: Tensor) -> Tensor:
        return obj


class qProbabilityOfImprovement(qLogExpectedImprovement):
    """MC-based batch Probability of Improvement: mean_S max_q sigmoid((y - best_f) / tau)."""

    _log = False

    def __init__(self, model, best_f, sampler=None, objective=None, posterior_transform=None, X_pending=None,
                 tau: float = 1e-3, constraints=None, eta=1e-3) -> None:
        super().__init__(model=model, best_f=best_f, sampler=sampler, objective=objective,
                         posterior_transform=posterior_transform, X_pending=X_pending, constraints=constraints, eta=eta,
                         tau_relu=tau)
        self.register_buffer("tau", torch.as_tensor(tau))
        _use_plain_reductions(self)
        self._utility_mode = 4

    def _sample_forward(self, obj: Tensor) -> Tensor:
        return torch.sigmoid((obj - self.best_f.unsqueeze(-1).to(obj)) / self.tau)


class qNoisyExpectedImprovement(qLogNoisyExpectedImprovement):
    """MC-based batch Noisy Expected Improvement: mean_S max_q relu(y - max_j y_baseline,j), cached-root sampling."""

    _log = False

    def __init__(self, model, X_baseline: Tensor, sampler=None, objective=None, posterior_transform=None, X_pending=None,
                 prune_baseline: bool = True, cache_root: bool = True, constraints=None, eta=1e-3,
                 marginalize_dim=None) -> None:
        super().__init__(model=model, X_baseline=X_baseline, sampler=sampler, objective=objective,
                         posterior_transform=posterior_transform, X_pending=X_pending, constraints=constraints, eta=eta,
                         prune_baseline=prune_baseline, cache_root=cache_root, marginalize_dim=marginalize_dim,
                         incremental=False)
        _use_plain_reductions(self)
        self._utility_mode = 2

    def _sample_forward(self, obj: Tensor) -> Tensor:
        return (obj - self.compute_best_f(obj).unsqueeze(-1)).clamp_min(0)


# ---- utilities whose per-sample value depends on the MC mean over ALL samples --------------------------------------------
# mu_i = mean_s obj[s][i] is linear in the base samples, mu_i = w (mean_i + sum_j coef_ij Zbar_j) + o with Zbar the sample
# means of the base samples, so the kernels evaluate it in closed form and stay single-pass (utility modes 5 / 6 of
# `csrc/sample_reduce.cu`); the torch `_sample_forward` below is the generic route (custom objectives, posterior transforms).
class qUpperConfidenceBound(qLogExpectedImprovement):
    """MC-based batch UCB: mean_S max_q (mu + sqrt(beta pi / 2) |y - mu|), mu = MC mean (reference :833-906)."""

    _log = False

    def __init__(self, model, beta: float, sampler=None, objective=None, posterior_transform=None, X_pending=None) -> None:
        super().__init__(model=model, best_f=0.0, sampler=sampler, objective=objective,
                         posterior_transform=posterior_transform, X_pending=X_pending)
        _use_plain_reductions(self)
        self.beta_prime = self._get_beta_prime(beta=beta)
        self._utility_mode = 5
        self._util_param = self.beta_prime

    def _get_beta_prime(self, beta: float) -> float:
        return math.sqrt(beta * math.pi / 2)

    def _sample_forward(self, obj: Tensor) -> Tensor:
        mean = obj.mean(dim=0)
        return mean + self.beta_prime * (obj - mean).abs()


class qLowerConfidenceBound(qUpperConfidenceBound):
    """Pessimistic counterpart of qUCB (reference :909-921)."""

    def _get_beta_prime(self, beta: float) -> float:
        return -super()._get_beta_prime(beta=beta)


class qPosteriorStandardDeviation(qLogExpectedImprovement):
    """MC-based batch posterior standard deviation: mean_S max_q sqrt(pi / 2) |y - mu| (reference :924-989)."""

    _log = False

    def __init__(self, model, sampler=None, objective=None, posterior_transform=None, X_pending=None, constraints=None,
                 eta=1e-3) -> None:
        super().__init__(model=model, best_f=0.0, sampler=sampler, objective=objective,
                         posterior_transform=posterior_transform, X_pending=X_pending, constraints=constraints, eta=eta)
        _use_plain_reductions(self)
        self._scale = math.sqrt(math.pi / 2)
        self._utility_mode = 6
        self._util_param = self._scale

    def _sample_forward(self, obj: Tensor) -> Tensor:
        mean = obj.mean(dim=0)
        return (obj - mean).abs() * self._scale
