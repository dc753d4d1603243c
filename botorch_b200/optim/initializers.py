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
from ..exceptions.warnings import BadInitialCandidatesWarning, SamplingWarning
from ..utils.sampling import boltzmann_sample, draw_sobol_samples, manual_seed
from .sharded import sharded_evaluate


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
    seed = options.get("seed")
    batch_limit = options.get("init_batch_limit", options.get("batch_limit"))
    device = bounds.device
    bounds_cpu = bounds.cpu()
    if options.get("topn"):
        init_func, opts = initialize_q_batch_topn, ("sorted", "largest")
    else:
        init_func, opts = initialize_q_batch, ("eta",)
    init_kwargs = {o: options[o] for o in opts if o in options}
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
