"""Setup-time helpers of qLogNEI (reference: botorch/acquisition/utils.py:97-178, 347-437)."""
from __future__ import annotations

import math
import warnings

import torch
from torch import Tensor

from ..exceptions.errors import UnsupportedError
from ..exceptions.warnings import BotorchWarning
from ..sampling.get_sampler import get_sampler
from ..utils.objective import compute_feasibility_indicator, compute_smoothed_feasibility_indicator


def compute_best_feasible_objective(samples: Tensor, obj: Tensor, constraints=None, model=None, objective=None,
                                    posterior_transform=None, X_baseline: Tensor | None = None,
                                    infeasible_obj: Tensor | None = None) -> Tensor:
    """Largest objective among the FEASIBLE points per sample (reference :97-178); without constraints `obj.amax(-1)`.
    With no feasible point, `infeasible_obj` or a 6-sigma lower bound of the objective over the box of `X_baseline`."""
    if constraints is None:
        with torch.no_grad():
            return obj.amax(dim=-1, keepdim=False)
    is_feasible = compute_feasibility_indicator(constraints=constraints, samples=samples)
    if is_feasible.any(dim=-1).all():
        infeasible_value = -torch.inf
    elif infeasible_obj is not None:
        infeasible_value = infeasible_obj.item()
    else:
        if model is None:
            raise ValueError("Must specify `model` when no feasible observation exists.")
        if X_baseline is None:
            raise ValueError("Must specify `X_baseline` when no feasible observation exists.")
        warnings.warn("When all training points are infeasible, it is better to use q(Log)ProbabilityOfFeasibility.",
                      BotorchWarning, stacklevel=2)
        infeasible_value = _estimate_objective_lower_bound(model=model, objective=objective,
                                                           posterior_transform=posterior_transform, X=X_baseline).item()
    while is_feasible.ndim < obj.ndim:  # augmented objectives carry extra leading dims
        is_feasible = is_feasible.unsqueeze(0)
    obj = torch.where(is_feasible.expand_as(obj), obj, infeasible_value)
    with torch.no_grad():
        return obj.amax(dim=-1, keepdim=False)


def get_infeasible_cost(X: Tensor, model, objective=None, posterior_transform=None) -> Tensor:
    """`M >= -min_x objective(x)` from the 6-sigma band of the posterior at `X` (reference :222-273)."""
    if objective is None:
        def objective(Y: Tensor, X: Tensor | None = None):
            return Y.squeeze(-1)

    posterior = model.posterior(X, posterior_transform=posterior_transform)
    six_stdv = 6 * posterior.variance.clamp_min(0).sqrt()
    lb = torch.stack([objective(posterior.mean - six_stdv, X=X), objective(posterior.mean + six_stdv, X=X)], dim=0)
    lb = lb.min(dim=0).values
    if lb.ndim < posterior.mean.ndim:
        lb = lb.unsqueeze(-1)
    while lb.dim() > 1:
        lb = lb.min(dim=-2).values
    return -(lb.clamp_max(0.0))


def _estimate_objective_lower_bound(model, objective, posterior_transform, X: Tensor) -> Tensor:
    """-M at 32 uniform points of the 10 %-padded bounding box of `X` (reference :181-219)."""
    X_lb = X.min(dim=-2, keepdim=True).values
    X_ub = X.max(dim=-2, keepdim=True).values
    X_range = X_ub - X_lb
    X_padding = 0.1 * X_range
    uniform = torch.rand(*X.shape[:-2], 32, X.shape[-1], dtype=X.dtype, device=X.device)
    X_samples = X_lb - X_padding + uniform * (X_range + 2 * X_padding)
    return -get_infeasible_cost(X=X_samples, model=model, objective=objective, posterior_transform=posterior_transform)


def prune_inferior_points(model, X: Tensor, objective=None, posterior_transform=None, constraints=None,
                          num_samples: int = 2048, max_frac: float = 1.0, sampler=None, marginalize_dim=None) -> Tensor:
    """Keep the points of `X` that are the arg-max of at least one joint posterior sample (:347-437); infeasible points
    rank below every feasible one, and if nothing is feasible the log-feasibility takes the objective's place."""
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
    infeas = ~compute_feasibility_indicator(constraints=constraints, samples=samples, marginalize_dim=marginalize_dim)
    if infeas.all():
        obj_vals = compute_smoothed_feasibility_indicator(constraints=constraints, samples=samples, eta=1e-3, log=True)
    elif infeas.any():
        obj_vals = obj_vals.clone()
        obj_vals[infeas] = obj_vals.min() - 1
    is_best = torch.argmax(obj_vals, dim=-1)
    idcs, counts = torch.unique(is_best, return_counts=True)
    if len(idcs) > max_points:
        counts, order_idcs = torch.sort(counts, stable=True, descending=True)
        idcs = order_idcs[:max_points]
    return X[idcs]
