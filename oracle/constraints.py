"""Oracle for outcome constraints (TEST INFRASTRUCTURE ONLY): CPU restatement of
botorch/utils/objective.py:98-211 (feasibility indicators), botorch/utils/safe_math.py:84-98, 434-458, 493-507
(log1pexp / logexpit / fatmoid / sigmoid) and of the constraint weighting in
botorch/acquisition/monte_carlo.py:322-348, applied to qLogEI / qEI on independent-output GPs
(models/gpytorch.py:786-884).  Pinned against reference-generated vectors: tests/golden/constraints_ref.pt
(tests/golden/make_golden_constraints.py)."""
from __future__ import annotations

import math

import torch
from torch import Tensor

from . import safe_math as sm
from .gp import OracleGP, mvn_rsample_from_base_samples
from .sampling import draw_sobol_normal_samples


def log1pexp(x: Tensor) -> Tensor:
    out = torch.empty_like(x)
    lo = x <= 18
    out[lo] = torch.log1p(torch.exp(x[lo]))
    out[~lo] = x[~lo] + torch.exp(-x[~lo])
    return out


def logexpit(x: Tensor) -> Tensor:
    return -log1pexp(-x)


def fatmoid(x: Tensor, tau=1.0) -> Tensor:
    u = x / tau
    m = math.sqrt(1 / 3)
    neg = (2 / 3) / (1 + (u - m) ** 2)
    pos = 1 - (2 / 3) / (1 + (u + m) ** 2)
    return torch.where(u < 0, neg, pos)


def feasibility_indicator(constraints, samples: Tensor) -> Tensor:
    ok = torch.ones(samples.shape[:-1], dtype=torch.bool)
    for c in constraints or []:
        ok = ok & (c(samples) <= 0)
    return ok


def smoothed_feasibility(constraints, samples: Tensor, eta, log: bool, fat) -> Tensor:
    etas = eta if isinstance(eta, Tensor) else torch.full((len(constraints),), float(eta))
    fats = fat if isinstance(fat, list) else [fat] * len(constraints)
    acc = torch.zeros_like(samples[..., 0])
    for c, e, f in zip(constraints, etas, fats):
        if f is None:
            acc = acc + torch.log(c(samples))
        elif f:
            acc = acc + torch.log(fatmoid(-c(samples) / e))
        else:
            acc = acc + logexpit(-c(samples) / e)
    return acc if log else torch.exp(acc)


class OracleConstrainedQEI:
    """qLogEI (log=True) or qEI (log=False) with outcome constraints on an m-output independent model: objective =
    output 0, samples drawn output by output from the S x 1 x q x m Sobol base samples (column k <-> output k)."""

    def __init__(self, gps: list[OracleGP], constraints, best_f, S: int, seed: int, log: bool = True, eta=1e-3,
                 tau_relu: float = sm.TAU_RELU, tau_max: float = sm.TAU_MAX, fat: bool = True) -> None:
        self.gps, self.constraints, self.S, self.seed, self.log, self.eta = gps, constraints, S, seed, log, eta
        self.best_f = torch.as_tensor(best_f, dtype=torch.float64)
        self.tau_relu, self.tau_max, self.fat = tau_relu, tau_max, fat

    def samples(self, X: Tensor) -> Tensor:
        b, q, _ = X.shape
        m = len(self.gps)
        Z = draw_sobol_normal_samples(q * m, self.S, X.dtype, self.seed).view(self.S, 1, q, m)
        outs = []
        for k, gp in enumerate(self.gps):
            mean, cov = gp.posterior_mvn(X)
            outs.append(mvn_rsample_from_base_samples(mean, cov, Z[..., k].expand(self.S, b, q), (self.S,)).squeeze(-1))
        return torch.stack(outs, dim=-1)  # S x b x q x m

    def __call__(self, X: Tensor) -> Tensor:
        if X.dim() == 2:
            X = X.unsqueeze(0)
        Y = self.samples(X)
        obj = Y[..., 0]
        if self.log:
            util = sm.log_improvement(obj, self.best_f, self.tau_relu, self.fat)
            util = util + smoothed_feasibility(self.constraints, Y, self.eta, log=True, fat=self.fat)
            red = sm.fatmax(util, dim=-1, tau=self.tau_max) if self.fat else sm.smooth_amax(util, dim=-1, tau=self.tau_max)
            return sm.logmeanexp(red, dim=0)
        util = (obj - self.best_f).clamp_min(0) * smoothed_feasibility(self.constraints, Y, self.eta, log=False, fat=False)
        return util.amax(dim=-1).mean(dim=0)
