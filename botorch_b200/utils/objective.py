"""Outcome-constraint indicators on posterior samples (reference: botorch/utils/objective.py:98-211).

Host-side torch ops: constraints are arbitrary user callables on `sample_shape x b x q x m` samples, so this part of a
constrained acquisition function runs on the generic (unfused) sample-reducing route; the samples themselves still come
from the CUDA posterior kernels."""
from __future__ import annotations

from typing import Callable

import torch
from torch import Tensor

from .safe_math import log_fatmoid, logexpit


def compute_feasibility_indicator(constraints: list[Callable[[Tensor], Tensor]] | None, samples: Tensor,
                                  marginalize_dim: int | None = None) -> Tensor:
    """Boolean `... x q` indicator: every constraint value <= 0 (reference :98-132)."""
    ind = torch.ones(samples.shape[:-1], dtype=torch.bool, device=samples.device)
    if constraints is not None:
        for con in constraints:
            ind = ind.logical_and(con(samples) <= 0)
    if ind.ndim >= 3 and marginalize_dim is not None:
        if marginalize_dim < 0:
            marginalize_dim = 1 + marginalize_dim % ind.ndim  # the output dim is already gone (reference :126-129)
        ind = ind.float().mean(dim=marginalize_dim).round().bool()
    return ind


def compute_smoothed_feasibility_indicator(constraints: list[Callable[[Tensor], Tensor]], samples: Tensor,
                                           eta: Tensor | float, log: bool = False,
                                           fat: list[bool | None] | bool = False) -> Tensor:
    """Product (sum of logs) over constraints of sigmoid / fatmoid(-c(samples) / eta); `fat_i is None` takes the callable's
    value as a probability directly (reference :135-211)."""
    if type(eta) is not Tensor:
        eta = torch.full((len(constraints),), eta)
    if type(fat) is not list:
        fat = [fat] * len(constraints)
    if len(eta) != len(constraints):
        raise ValueError("Number of provided constraints and number of provided etas do not match.")
    if len(fat) != len(constraints):
        raise ValueError("Number of provided constraints and number of provided fats do not match.")
    if not (eta > 0).all():
        raise ValueError("eta must be positive.")
    log_feas = torch.zeros_like(samples[..., 0])
    for con, eta_i, fat_i in zip(constraints, eta, fat):
        if fat_i is None:
            log_feas = log_feas + con(samples).log()
        else:
            log_feas = log_feas + (log_fatmoid if fat_i else logexpit)(-con(samples) / eta_i)
    return log_feas if log else log_feas.exp()
