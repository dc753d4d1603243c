"""Initial-condition generation for multi-start optimisation (reference: botorch/optim/initializers.py:258-483,
959-1172).  `X_rnd` is drawn and kept on the host exactly like the reference; the raw-sample sweep is the
biggest batched forward of the whole optimisation (b = raw_samples) and can be sharded across ranks."""
from __future__ import annotations

import warnings

import torch
import torch.distributed as dist
from torch import Tensor
from torch.quasirandom import SobolEngine

from ..exceptions.errors import UnsupportedError
from ..exceptions.warnings import BadInitialCandidatesWarning, BotorchWarning, SamplingWarning
from ..utils.sampling import (boltzmann_sample, draw_sobol_samples, manual_seed, sample_perturbed_subset_dims,
                              sample_truncated_normal_perturbations)
from .sharded import sharded_evaluate
from .utils import get_X_baseline


def initialize_q_batch(X: Tensor, acq_vals: Tensor, n: int, eta: float = 1.0) -> tuple[Tensor, Tensor]:
    """Boltzmann selection without replacement + forced inclusion of the arg-max (reference :959-1042)."""
    n_samples = X.shape[0]
    if n > n_samples:
        raise RuntimeError(f"n ({n}) cannot be larger than the number of provided samples ({n_samples})")
    if n == n_samples:
        return X, acq_vals
    if X.dim() != 3:
        raise UnsupportedError("botorch_b200.initialize_q_batch supports un-batched `b x q x d` samples.")
    Ystd = acq_vals.std(dim=0)
    if torch.any(Ystd == 0) or not torch.isfinite(Ystd).all():
        warnings.warn("All acquisition values for raw samples points are the same or not finite for at least one "
                      "batch. Choosing initial conditions at random.", BadInitialCandidatesWarning, stacklevel=3)
        idcs = torch.randperm(n=n_samples, device=X.device)[:n]
        return X[idcs], acq_vals[idcs]
    max_val, max_idx = torch.max(acq_vals, dim=0)
    idcs = boltzmann_sample(acq_vals, num_samples=n, eta=eta)
    if max_idx not in idcs:
        idcs[-1] = max_idx
    return X[idcs], acq_vals[idcs]


def initialize_q_batch_nonneg(X: Tensor, acq_vals: Tensor, n: int, eta: float = 1.0, alpha: float = 1e-4):
    """Selection heuristic for non-negative acquisition functions that vanish over large areas (qEI & co.): ignore samples
    below `alpha * max`, weight the rest by exp(eta (v / max - 1)) (reference :1045-1121)."""
    n_samples = X.shape[0]
    if n > n_samples:
        raise RuntimeError("n cannot be larger than the number of provided samples")
    if n == n_samples:
        return X, acq_vals
    max_val, max_idx = torch.max(acq_vals, dim=0)
    if torch.any(max_val <= 0):
        warnings.warn("All acquisition values for raw sampled points are nonpositive, so initial conditions are being "
                      "selected randomly.", BadInitialCandidatesWarning, stacklevel=3)
        idcs = torch.randperm(n=n_samples, device=X.device)[:n]
        return X[idcs], acq_vals[idcs]
    pos = acq_vals > 0
    num_pos = pos.sum().item()
    if num_pos < n:
        # all positive points, the remaining quota filled with randomly selected ones
        remaining_indices = (~pos).nonzero(as_tuple=False).view(-1)
        rand_indices = torch.randperm(remaining_indices.shape[0], device=acq_vals.device)
        pos[remaining_indices[rand_indices[: n - num_pos]]] = 1
        return X[pos], acq_vals[pos]
    alpha_pos = acq_vals >= alpha * max_val
    while alpha_pos.sum() < n:
        alpha = 0.1 * alpha
        alpha_pos = acq_vals >= alpha * max_val
    alpha_pos_idcs = torch.arange(len(acq_vals), device=acq_vals.device)[alpha_pos]
    weights = torch.exp(eta * (acq_vals[alpha_pos] / max_val - 1))
    idcs = alpha_pos_idcs[torch.multinomial(weights, n)]
    if max_idx not in idcs:
        idcs[-1] = max_idx
    return X[idcs], acq_vals[idcs]


def is_nonnegative(acq_function) -> bool:
    """True for the acquisition classes known to be non-negative (reference :1281-1309; the analytic and hypervolume
    members of that list are outside this package)."""
    from ..acquisition.mc_improvement import (qExpectedImprovement, qNoisyExpectedImprovement,
                                              qProbabilityOfImprovement)

    return isinstance(acq_function, (qExpectedImprovement, qNoisyExpectedImprovement, qProbabilityOfImprovement))


def sample_points_around_best(acq_function, n_discrete_points: int, sigma: float, bounds: Tensor, best_pct: float = 5.0,
                              subset_sigma: float = 1e-1, prob_perturb: float | None = None) -> Tensor | None:
    """Perturbations of the best `best_pct` percent of the baseline points by posterior-mean objective
    (reference :1175-1278; single-objective branch -- the Pareto branch belongs to the multi-objective package)."""
    X = get_X_baseline(acq_function=acq_function)
    if X is None:
        return None
    with torch.no_grad():
        try:
            posterior = acq_function.model.posterior(X)
        except AttributeError:
            warnings.warn("Failed to sample around previous best points.", BotorchWarning, stacklevel=3)
            return None
        mean = posterior.mean
        while mean.ndim > 2:
            mean = mean.mean(dim=0)
        try:
            f_pred = acq_function.objective(mean)
        except (AttributeError, TypeError):
            f_pred = mean
        if hasattr(acq_function, "maximize") and not acq_function.maximize:
            f_pred = -f_pred
        constraints = getattr(acq_function, "constraints", None)
        if constraints is not None:
            neg_violation = -torch.stack([c(mean).clamp_min(0.0) for c in constraints], dim=-1).sum(dim=-1)
            feas = neg_violation == 0
            if feas.any():
                f_pred[~feas] = float("-inf")
            else:
                f_pred = neg_violation
        if f_pred.ndim == mean.ndim and f_pred.shape[-1] > 1:
            raise UnsupportedError("sample_points_around_best: multi-objective predictions are not supported here.")
        if f_pred.shape[-1] == 1:
            f_pred = f_pred.squeeze(-1)
        n_best = max(1, round(X.shape[0] * best_pct / 100))
        best_X = X[torch.topk(f_pred, n_best).indices.view(-1)]
    use_perturbed_sampling = best_X.shape[-1] >= 20 or prob_perturb is not None
    n_trunc_normal_points = n_discrete_points // 2 if use_perturbed_sampling else n_discrete_points
    perturbed_X = sample_truncated_normal_perturbations(X=best_X, n_discrete_points=n_trunc_normal_points, sigma=sigma,
                                                        bounds=bounds)
    if use_perturbed_sampling:
        perturbed_subset_dims_X = sample_perturbed_subset_dims(
            X=best_X, bounds=bounds, n_discrete_points=n_discrete_points - n_trunc_normal_points, sigma=sigma,
            prob_perturb=prob_perturb)
        perturbed_X = torch.cat([perturbed_X, perturbed_subset_dims_X], dim=0)
        perturbed_X = perturbed_X[torch.randperm(perturbed_X.shape[0], device=X.device)]
    return perturbed_X


def initialize_q_batch_topn(X: Tensor, acq_vals: Tensor, n: int, largest: bool = True, sorted: bool = True):
    """Deterministic top-n selection (reference :1124-1172)."""
    n_samples = X.shape[0]
    if n > n_samples:
        raise RuntimeError(f"n ({n}) cannot be larger than the number of provided samples ({n_samples})")
    if n == n_samples:
        return X, acq_vals
    if torch.any(acq_vals.std(dim=0) == 0):
        warnings.warn("All acquisition values for raw samples points are the same for at least one batch. Choosing "
                      "initial conditions at random.", BadInitialCandidatesWarning, stacklevel=3)
        idcs = torch.randperm(n=n_samples, device=X.device)[:n]
        return X[idcs], acq_vals[idcs]
    topk_out, topk_idcs = acq_vals.topk(n, largest=largest, sorted=sorted)
    return X[topk_idcs], topk_out


def gen_batch_initial_conditions(acq_function, bounds: Tensor, q: int, num_restarts: int, raw_samples: int,
                                 fixed_features=None, options: dict | None = None, inequality_constraints=None,
                                 equality_constraints=None, generator=None, fixed_X_fantasies=None,
                                 shard_across_ranks: bool = False) -> Tensor:
    """`num_restarts x q x d` initial conditions chosen from `raw_samples` Sobol q-batches by acquisition value."""
    if inequality_constraints or equality_constraints or fixed_features or fixed_X_fantasies is not None:
        raise UnsupportedError("botorch_b200.gen_batch_initial_conditions supports box bounds only.")
    if bounds.isinf().any():
        raise NotImplementedError("Currently only finite values in `bounds` are supported for generating initial "
                                  "conditions for optimization.")
    options = options or {}
    sample_around_best = options.get("sample_around_best", False)
    if sample_around_best and generator:
        raise UnsupportedError("Option 'sample_around_best' is not supported when custom generator is be used.")
    seed = options.get("seed")
    sharded = shard_across_ranks and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    if sharded:
        seed = _consistent_sharded_seed(acq_function, seed, bounds.device)
    batch_limit = options.get("init_batch_limit", options.get("batch_limit"))
    device = bounds.device
    bounds_cpu = bounds.cpu()
    if options.get("topn"):
        init_func, opts = initialize_q_batch_topn, ("sorted", "largest")
    elif options.get("nonnegative") or is_nonnegative(acq_function):
        init_func, opts = initialize_q_batch_nonneg, ("alpha", "eta")
    else:
        init_func, opts = initialize_q_batch, ("eta",)
    init_kwargs = {}
    for o in opts:
        if o == "largest" and hasattr(acq_function, "maximize"):
            init_kwargs[o] = acq_function.maximize
        if o in options:
            init_kwargs[o] = options[o]
    q = 1 if q is None else q
    effective_dim = bounds.shape[-1] * q
    if effective_dim > SobolEngine.MAXDIM:
        warnings.warn(f"Sample dimension q*d={effective_dim} exceeding Sobol max dimension ({SobolEngine.MAXDIM}). "
                      "Using iid samples instead.", SamplingWarning, stacklevel=3)
    factor, max_factor = 1, 5
    batch_initial_conditions = None
    while factor < max_factor:
        with warnings.catch_warnings(record=True) as ws:
            warnings.simplefilter("always")
            n = raw_samples * factor
            on_device = False
            if generator is not None:
                X_rnd = generator(n, q, seed)
            elif effective_dim <= SobolEngine.MAXDIM:
                # with CUDA bounds the Sobol cloud is generated on the device (bit-identical to the host engine's draw):
                # no host draw of n*q*d doubles, no H2D copy of them
                on_device = device.type == "cuda"
                X_rnd = draw_sobol_samples(bounds=bounds if on_device else bounds_cpu, n=n, q=q, seed=seed)
            else:
                with manual_seed(seed):
                    X_nlzd = torch.rand(n, q, bounds_cpu.shape[-1], dtype=bounds.dtype)
                X_rnd = X_nlzd * (bounds_cpu[1] - bounds_cpu[0]) + bounds_cpu[0]
            if sample_around_best:
                # reference :413-431: n more q-batches drawn around the best baseline points
                X_best_rnd = sample_points_around_best(
                    acq_function=acq_function, n_discrete_points=n * q, sigma=options.get("sample_around_best_sigma", 1e-3),
                    bounds=bounds, subset_sigma=options.get("sample_around_best_subset_sigma", 1e-1),
                    prob_perturb=options.get("sample_around_best_prob_perturb"))
                if X_best_rnd is not None:
                    X_best_rnd = X_best_rnd.view(n, q, bounds.shape[-1])
                    if sharded:  # the perturbations consume each process's own RNG: every rank adopts rank 0's
                        X_best_rnd = X_best_rnd.to(device).contiguous()
                        dist.broadcast(X_best_rnd, src=0)
                    X_rnd = torch.cat([X_rnd, X_best_rnd.to(X_rnd)], dim=0)
            if not on_device:
                X_rnd = X_rnd.cpu()
            with torch.no_grad():
                limit = X_rnd.shape[0] if batch_limit is None else batch_limit
                if shard_across_ranks:
                    acq_vals = sharded_evaluate(
                        lambda Xs: torch.cat([acq_function(x_.to(device=device)) for x_ in Xs.split(limit, dim=0)]),
                        X_rnd).cpu()
                else:
                    acq_vals = torch.cat([acq_function(x_.to(device=device)).cpu() for x_ in X_rnd.split(limit, dim=0)])
            if on_device:
                # the selection (Boltzmann draw from the host RNG, arg-max inclusion) runs on the host exactly as in the
                # reference, on the values and a row-index proxy; the chosen q-batches are then gathered on the device
                proxy = torch.arange(X_rnd.shape[0], dtype=torch.float64).view(-1, 1, 1)
                picked, _ = init_func(X=proxy, acq_vals=acq_vals, n=num_restarts, **init_kwargs)
                batch_initial_conditions = X_rnd[picked.reshape(-1).long().to(device)]
            else:
                batch_initial_conditions, _ = init_func(X=X_rnd, acq_vals=acq_vals, n=num_restarts, **init_kwargs)
            batch_initial_conditions = batch_initial_conditions.to(device=device)
            if shard_across_ranks and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                # the Boltzmann draw of `initialize_q_batch` consumes each process's own global RNG: every rank adopts
                # rank 0's selection so that the restarts the ranks split are slices of ONE list
                dist.broadcast(batch_initial_conditions, src=0)
            if not any(issubclass(w.category, BadInitialCandidatesWarning) for w in ws):
                return batch_initial_conditions
            if factor < max_factor:
                factor += 1
                if seed is not None:
                    seed += 1
    warnings.warn("Unable to find non-zero acquisition function values - initial conditions are being selected "
                  "randomly.", BadInitialCandidatesWarning, stacklevel=2)
    return batch_initial_conditions


def _consistent_sharded_seed(acq_function, seed: int | None, device) -> int:
    """Sharded sweeps all-gather values computed by different ranks, which is only meaningful when every rank evaluates the
    SAME point cloud with the SAME base samples.  An unset Sobol seed is drawn on rank 0 and broadcast; a sampler whose seed
    (hence base samples, and for qLogNEI the baseline draws behind `best_f`) differs between ranks is an error."""
    dev = device if device.type == "cuda" else torch.device("cpu")
    box = torch.tensor([seed if seed is not None else int(torch.randint(0, 2**31 - 1, (1,)).item())], dtype=torch.int64,
                       device=dev)
    dist.broadcast(box, src=0)
    sampler = getattr(acq_function, "sampler", None)
    sseed = torch.tensor([int(getattr(sampler, "seed", -1)) if sampler is not None else -1], dtype=torch.int64, device=dev)
    lo, hi = sseed.clone(), sseed.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if int(lo) != int(hi):
        raise RuntimeError("shard_across_ranks=True needs the same MC sampler seed on every rank (pass an explicit `seed` to "
                           "the sampler, or construct the acquisition function after seeding torch identically).")
    return int(box)
