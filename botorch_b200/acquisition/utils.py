"""Setup-time helpers of qLogNEI (reference: botorch/acquisition/utils.py:97-178, 347-437)."""
from __future__ import annotations

import math

import torch
from torch import Tensor

from ..exceptions.errors import UnsupportedError
from ..sampling.get_sampler import get_sampler


def compute_best_feasible_objective(samples: Tensor, obj: Tensor, constraints=None, **kwargs) -> Tensor:
    """Unconstrained case of the reference (:134-138): `obj.amax(-1)` without gradient."""
    if constraints is not None:
        raise UnsupportedError("Outcome constraints are on the 'next' list of botorch_b200 (SURVEY.md section 8f N3).")
    with torch.no_grad():
        return obj.amax(dim=-1, keepdim=False)


def prune_inferior_points(model, X: Tensor, objective=None, posterior_transform=None, constraints=None,
                          num_samples: int = 2048, max_frac: float = 1.0, sampler=None, marginalize_dim=None) -> Tensor:
    """Keep the points of `X` that are the arg-max of at least one joint posterior sample (:347-437)."""
    if constraints is not None:
        raise UnsupportedError("Outcome constraints are on the 'next' list of botorch_b200 (SURVEY.md section 8f N3).")
    if X.ndim > 2:
        raise UnsupportedError("Batched inputs `X` are currently unsupported by `prune_inferior_points`")
    if X.size(-2) == 0:
        raise ValueError("X must have at least one point.")
    if max_frac <= 0 or max_frac > 1.0:
        raise ValueError(f"max_frac must take values in (0, 1], is {max_frac}")
    max_points = math.ceil(max_frac * X.size(-2))
    with torch.no_grad():
        posterior = model.posterior(X=X, posterior_transform=posterior_transform)
    if sampler is None:
        sampler = get_sampler(posterior=posterior, sample_shape=torch.Size([num_samples]))
    samples = sampler(posterior)
    obj_vals = objective(samples=samples, X=X) if objective is not None else samples.squeeze(-1)
    if obj_vals.ndim > 2:
        raise UnsupportedError("Models with multiple batch dims are currently unsupported by `prune_inferior_points`.")
    is_best = torch.argmax(obj_vals, dim=-1)
    idcs, counts = torch.unique(is_best, return_counts=True)
    if len(idcs) > max_points:
        counts, order_idcs = torch.sort(counts, stable=True, descending=True)
        idcs = order_idcs[:max_points]
    return X[idcs]
