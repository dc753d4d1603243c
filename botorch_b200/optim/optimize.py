"""`optimize_acqf` (reference: botorch/optim/optimize.py:364-845): Sobol raw samples -> acquisition sweep ->
initial-condition selection -> batched L-BFGS-B -> arg-max.  With `shard_across_ranks=True` and an initialised
`torch.distributed` group the raw-sample sweep and the restarts are split across ranks (one process per GPU); the
only collective is the all-gather of acquisition values / candidates."""
from __future__ import annotations

import time
import warnings

import torch
import torch.distributed as dist
from torch import Tensor

from .. import settings
from ..exceptions.errors import UnsupportedError
from ..exceptions.warnings import OptimizationWarning
from .initializers import gen_batch_initial_conditions
from .sharded import shard_bounds

INIT_OPTION_KEYS = {"alpha", "batch_limit", "eta", "init_batch_limit", "nonnegative", "n_burnin", "sample_around_best",
                    "sample_around_best_sigma", "sample_around_best_subset_sigma", "sample_around_best_prob_perturb", "seed",
                    "thinning", "topn", "sorted", "largest"}


def _gather_cat(t: Tensor, sizes: list[int]) -> Tensor:
    """All-gather ragged first-dimension shards (equal trailing shape) and concatenate them in rank order."""
    world = dist.get_world_size()
    mx = max(sizes)
    pad = torch.zeros((mx, *t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[:s] for o, s in zip(out, sizes)])


def _combine_initial_conditions(provided: Tensor | None, generated: Tensor | None) -> Tensor:
    """reference :342-361: user-provided initial conditions first, generated ones appended."""
    if provided is not None and generated is not None:
        return torch.cat([provided, generated.to(provided)], dim=-3)
    if provided is not None:
        return provided
    if generated is not None:
        return generated
    raise ValueError("Either `batch_initial_conditions` or `raw_samples` must be set.")


def _optimize_acqf_batch(acq_function, bounds: Tensor, q: int, num_restarts: int, raw_samples: int | None,
                         options: dict, batch_initial_conditions: Tensor | None, return_best_only: bool,
                         shard_across_ranks: bool, post_processing_func=None, timeout_sec: float | None = None,
                         retry_on_optimization_warning: bool = True, gen_candidates=None, ic_generator=None):
    """reference :364-620 (box-bounded part: no feasibility projection)."""
    from ..generation.gen import gen_candidates_scipy  # local import: generation <-> optim are mutually dependent

    if gen_candidates is None:
        gen_candidates = gen_candidates_scipy
        if settings.optimizer.value() == "device" and bounds.is_cuda:
            from ..generation.device_gen import gen_candidates_device

            gen_candidates = gen_candidates_device
    ic_generator = gen_batch_initial_conditions if ic_generator is None else ic_generator
    sharded = shard_across_ranks and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    provided = batch_initial_conditions
    required_num_restarts = num_restarts
    if provided is not None and provided.ndim == 3:
        required_num_restarts -= provided.shape[0]

    def generate():
        if raw_samples is None or required_num_restarts <= 0:
            return None
        kw = dict(acq_function=acq_function, bounds=bounds, q=q, num_restarts=required_num_restarts,
                  raw_samples=raw_samples, options=options)
        if ic_generator is gen_batch_initial_conditions:
            kw["shard_across_ranks"] = sharded
        return ic_generator(**kw)

    ics_all = _combine_initial_conditions(provided, generate())
    batch_limit = options.get("batch_limit", num_restarts)
    lower = None if bounds[0].isinf().all() else bounds[0]
    upper = None if bounds[1].isinf().all() else bounds[1]
    gen_options = {k: v for k, v in options.items() if k not in INIT_OPTION_KEYS}

    def run(ics: Tensor):
        sizes = None
        if sharded:
            rank, world = dist.get_rank(), dist.get_world_size()
            sizes = [shard_bounds(ics.shape[0], r, world)[1] - shard_bounds(ics.shape[0], r, world)[0] for r in range(world)]
            lo, hi = shard_bounds(ics.shape[0], rank, world)
            ics = ics[lo:hi]
        batched = ics.split(batch_limit)
        per_batch_timeout = timeout_sec / max(1, len(batched)) if timeout_sec is not None else None
        cands, vals, ws_all = [], [], []
        for ics_ in batched:
            with warnings.catch_warnings(record=True) as ws:
                warnings.simplefilter("always", category=OptimizationWarning)
                c, v = gen_candidates(ics_, acq_function, lower_bounds=lower, upper_bounds=upper, options=gen_options,
                                      timeout_sec=per_batch_timeout)
            ws_all += ws
            cands.append(c)
            vals.append(v.reshape(-1))
        bc = torch.cat(cands) if cands else ics.new_zeros((0, *ics.shape[1:]))
        bv = torch.cat(vals) if vals else ics.new_zeros(0)
        if sharded:
            bc, bv = _gather_cat(bc, sizes), _gather_cat(bv, sizes)
        return bc, bv, ws_all

    batch_candidates, batch_acq_values, ws = run(ics_all)
    for w in ws:  # surface what was recorded (the reference's recorded warnings are summarised in the retry message)
        if not issubclass(w.category, OptimizationWarning):
            warnings.warn_explicit(w.message, w.category, w.filename, w.lineno)
    failed = any(issubclass(w.category, OptimizationWarning) for w in ws)
    if sharded and retry_on_optimization_warning:
        flag = torch.tensor([int(failed)], device=bounds.device if bounds.is_cuda else "cpu")
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)  # every rank must take the same branch
        failed = bool(flag.item())
    if failed and retry_on_optimization_warning:
        msgs = [w.message for w in ws]
        if provided is not None and required_num_restarts <= 0:
            warnings.warn("Optimization failed in `gen_candidates_scipy` with the following "
                          f"warning(s):\n{msgs}\nBecause you specified `batch_initial_conditions` larger than required "
                          "`num_restarts`, optimization will not be retried with new initial conditions and will proceed "
                          "with the current solution. Suggested remediation: Try again with different "
                          "`batch_initial_conditions`, don't provide `batch_initial_conditions`, or increase `num_restarts`.",
                          RuntimeWarning, stacklevel=3)
        else:
            warnings.warn("Optimization failed in `gen_candidates_scipy` with the following "
                          f"warning(s):\n{msgs}\nTrying again with a new set of initial conditions.", RuntimeWarning,
                          stacklevel=3)
        if raw_samples is not None and required_num_restarts > 0:
            ics_all = _combine_initial_conditions(provided, generate())
            batch_candidates, batch_acq_values, ws = run(ics_all)
            if any(issubclass(w.category, OptimizationWarning) for w in ws):
                warnings.warn("Optimization failed on the second try, after generating a new set of initial conditions.",
                              RuntimeWarning, stacklevel=3)
    elif failed:
        for w in ws:
            if issubclass(w.category, OptimizationWarning):
                warnings.warn_explicit(w.message, w.category, w.filename, w.lineno)

    if post_processing_func is not None:
        # reference :528-537: ALL restart candidates are post-processed and re-evaluated before the arg-max
        batch_candidates = post_processing_func(batch_candidates)
        with torch.no_grad():
            batch_acq_values = torch.cat([acq_function(c).reshape(-1) for c in batch_candidates.split(batch_limit, dim=0)])
    if return_best_only:
        best = torch.argmax(batch_acq_values.view(-1), dim=0)
        return batch_candidates[best], batch_acq_values[best]
    return batch_candidates, batch_acq_values


def optimize_acqf(acq_function, bounds: Tensor, q: int, num_restarts: int, raw_samples: int | None = None,
                  options: dict | None = None, inequality_constraints=None, equality_constraints=None,
                  nonlinear_inequality_constraints=None, fixed_features=None, post_processing_func=None,
                  batch_initial_conditions: Tensor | None = None, return_best_only: bool = True,
                  gen_candidates=None, sequential: bool = False, ic_generator=None, timeout_sec: float | None = None,
                  retry_on_optimization_warning: bool = True, shard_across_ranks: bool = False, **unused):
    """Multi-start optimisation of an acquisition function over a box.  Returns (candidates, acquisition value)."""
    if inequality_constraints or equality_constraints or nonlinear_inequality_constraints or fixed_features:
        raise UnsupportedError("botorch_b200.optimize_acqf supports box-bounded problems only.")
    if bounds.ndim != 2 or bounds.shape[0] != 2:
        raise ValueError(f"bounds should be a `2 x d` tensor, current shape: {list(bounds.shape)}.")
    if batch_initial_conditions is None and raw_samples is None and ic_generator is None:
        raise ValueError("Must specify `raw_samples` when `batch_initial_conditions` is None`.")
    options = options or {}
    common = dict(shard_across_ranks=shard_across_ranks, post_processing_func=post_processing_func,
                  retry_on_optimization_warning=retry_on_optimization_warning, gen_candidates=gen_candidates,
                  ic_generator=ic_generator)
    if sequential and q > 1:
        # sequential greedy (reference :232-339): q rounds with q=1 and growing X_pending
        if not return_best_only:
            raise NotImplementedError("`return_best_only=False` only supported for joint optimization.")
        if batch_initial_conditions is not None:
            raise UnsupportedError("`batch_initial_conditions` is not supported for sequential optimization. Either avoid "
                                   "specifying `batch_initial_conditions` to use the custom initializer or use the "
                                   "`ic_generator` kwarg to generate initial conditions for the case of "
                                   "nonlinear inequality constraints.")
        per_step_timeout = timeout_sec / q if timeout_sec is not None else None
        start = time.monotonic()
        candidate_list, acq_value_list = [], []
        base_X_pending = acq_function.X_pending
        candidates = torch.empty(0, bounds.shape[-1], dtype=bounds.dtype, device=bounds.device)
        for i in range(q):
            cand, val = _optimize_acqf_batch(acq_function, bounds, 1, num_restarts, raw_samples, options, None, True,
                                             timeout_sec=per_step_timeout, **common)
            candidate_list.append(cand)
            acq_value_list.append(val)
            candidates = torch.cat(candidate_list, dim=-2)
            acq_function.set_X_pending(torch.cat([base_X_pending, candidates], dim=-2)
                                       if base_X_pending is not None else candidates)
            if per_step_timeout is not None:  # re-allocate what is left of the budget (reference :325-329)
                per_step_timeout = max(timeout_sec - (time.monotonic() - start), 1e-6) / max(q - i - 1, 1)
        acq_function.set_X_pending(base_X_pending)
        return candidates, torch.stack(acq_value_list)
    return _optimize_acqf_batch(acq_function, bounds, q, num_restarts, raw_samples, options, batch_initial_conditions,
                                return_best_only, timeout_sec=timeout_sec, **common)
