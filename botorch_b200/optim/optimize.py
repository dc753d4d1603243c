"""`optimize_acqf` (reference: botorch/optim/optimize.py:364-845): Sobol raw samples -> acquisition sweep ->
initial-condition selection -> batched L-BFGS-B -> arg-max.  With `shard_across_ranks=True` and an initialised
`torch.distributed` group the raw-sample sweep and the restarts are split across ranks (one process per GPU); the
only collective is the all-gather of acquisition values / candidates."""
from __future__ import annotations

import warnings

import torch
import torch.distributed as dist
from torch import Tensor

from ..exceptions.errors import UnsupportedError
from .initializers import gen_batch_initial_conditions
from .sharded import shard_bounds

INIT_OPTION_KEYS = {"alpha", "batch_limit", "eta", "init_batch_limit", "nonnegative", "n_burnin", "sample_around_best",
                    "sample_around_best_sigma", "sample_around_best_prob_perturb", "seed", "thinning", "topn", "sorted",
                    "largest"}


def _gather_cat(t: Tensor, sizes: list[int]) -> Tensor:
    """All-gather ragged first-dimension shards (equal trailing shape) and concatenate them in rank order."""
    world = dist.get_world_size()
    mx = max(sizes)
    pad = torch.zeros((mx, *t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[:s] for o, s in zip(out, sizes)])


def _optimize_acqf_batch(acq_function, bounds: Tensor, q: int, num_restarts: int, raw_samples: int | None,
                         options: dict, batch_initial_conditions: Tensor | None, return_best_only: bool,
                         shard_across_ranks: bool):
    from ..generation.gen import gen_candidates_scipy  # local import: generation <-> optim are mutually dependent

    sharded = shard_across_ranks and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    if batch_initial_conditions is None:
        if raw_samples is None:
            raise ValueError("Must specify `raw_samples` when `batch_initial_conditions` is None`.")
        batch_initial_conditions = gen_batch_initial_conditions(
            acq_function=acq_function, bounds=bounds, q=q, num_restarts=num_restarts, raw_samples=raw_samples,
            options=options, shard_across_ranks=sharded)
    batch_limit = options.get("batch_limit", num_restarts)
    lower = None if bounds[0].isinf().all() else bounds[0]
    upper = None if bounds[1].isinf().all() else bounds[1]
    gen_options = {k: v for k, v in options.items() if k not in INIT_OPTION_KEYS}
    ics = batch_initial_conditions
    if sharded:
        rank, world = dist.get_rank(), dist.get_world_size()
        sizes = [shard_bounds(ics.shape[0], r, world)[1] - shard_bounds(ics.shape[0], r, world)[0] for r in range(world)]
        lo, hi = shard_bounds(ics.shape[0], rank, world)
        ics = ics[lo:hi]
    cands, vals = [], []
    for ics_ in ics.split(batch_limit):
        c, v = gen_candidates_scipy(ics_, acq_function, lower_bounds=lower, upper_bounds=upper, options=gen_options)
        cands.append(c)
        vals.append(v.reshape(-1))
    batch_candidates = torch.cat(cands) if cands else ics.new_zeros((0, *ics.shape[1:]))
    batch_acq_values = torch.cat(vals) if vals else ics.new_zeros(0)
    if sharded:
        batch_candidates = _gather_cat(batch_candidates, sizes)
        batch_acq_values = _gather_cat(batch_acq_values, sizes)
    if return_best_only:
        best = torch.argmax(batch_acq_values.view(-1), dim=0)
        return batch_candidates[best], batch_acq_values[best]
    return batch_candidates, batch_acq_values


def optimize_acqf(acq_function, bounds: Tensor, q: int, num_restarts: int, raw_samples: int | None = None,
                  options: dict | None = None, inequality_constraints=None, equality_constraints=None,
                  nonlinear_inequality_constraints=None, fixed_features=None, post_processing_func=None,
                  batch_initial_conditions: Tensor | None = None, return_best_only: bool = True,
                  sequential: bool = False, shard_across_ranks: bool = False, **unused):
    """Multi-start optimisation of an acquisition function over a box.  Returns (candidates, acquisition value)."""
    if inequality_constraints or equality_constraints or nonlinear_inequality_constraints or fixed_features:
        raise UnsupportedError("botorch_b200.optimize_acqf supports box-bounded problems only.")
    if bounds.ndim != 2 or bounds.shape[0] != 2:
        raise ValueError(f"bounds should be a `2 x d` tensor, current shape: {list(bounds.shape)}.")
    options = options or {}
    if sequential and q > 1:
        if not return_best_only:
            raise NotImplementedError("`return_best_only=False` only supported for joint optimization.")
        # sequential greedy (reference :278-339): q rounds with q=1 and growing X_pending
        candidate_list, acq_value_list = [], []
        base_X_pending = acq_function.X_pending
        candidates = torch.empty(0, bounds.shape[-1], dtype=bounds.dtype, device=bounds.device)
        for _ in range(q):
            cand, val = _optimize_acqf_batch(acq_function, bounds, 1, num_restarts, raw_samples, options, None, True,
                                             shard_across_ranks)
            if post_processing_func is not None:
                cand = post_processing_func(cand)
            candidate_list.append(cand)
            acq_value_list.append(val)
            candidates = torch.cat(candidate_list, dim=-2)
            acq_function.set_X_pending(torch.cat([base_X_pending, candidates], dim=-2)
                                       if base_X_pending is not None else candidates)
        acq_function.set_X_pending(base_X_pending)
        return candidates, torch.stack(acq_value_list)
    cand, val = _optimize_acqf_batch(acq_function, bounds, q, num_restarts, raw_samples, options,
                                     batch_initial_conditions, return_best_only, shard_across_ranks)
    if post_processing_func is not None:
        cand = post_processing_func(cand)
        with torch.no_grad():
            val = acq_function(cand)
    return cand, val
