"""`gen_candidates_scipy` (reference: botorch/generation/gen.py:62-507), box-bounded fast path.

Each L-BFGS-B round moves `b' x q x d` doubles host->device and `b' (1 + q d)` doubles back
(gen.py:435-444, 469-485); value and gradient come from one fused CUDA forward+backward.
"""
from __future__ import annotations

import warnings
from functools import partial

import numpy as np
import torch
from torch import Tensor

from .. import settings

from ..exceptions.errors import OptimizationGradientError, UnsupportedError
from ..exceptions.warnings import OptimizationWarning
from ..optim.batched_lbfgs_b import fmin_l_bfgs_b_batched
from ..optim.utils import _arrayify, columnwise_clamp


def _f_np_wrapper(x: np.ndarray, f, shapeX, device, dtype, batch_indices=None):
    """numpy -> torch -> (losses, grad of losses.sum()) -> numpy (gen.py:423-485)."""
    if np.isnan(x).any():
        raise RuntimeError(f"{np.isnan(x).sum()} elements of the {x.size} element array `x` are NaN.")
    X = torch.from_numpy(x).to(device=device, dtype=dtype).view(-1, *shapeX[1:]).contiguous().requires_grad_(True)
    losses = f(X)
    loss = losses.sum()
    gradf = _arrayify(torch.autograd.grad(loss, X)[0].contiguous().view(-1)).reshape(*x.shape)
    if np.isnan(gradf).any():
        raise OptimizationGradientError(
            f"{np.isnan(gradf).sum()} elements of the {x.size} element gradient array `gradf` are NaN. This often "
            "indicates numerical issues.", current_x=x)
    fval = losses.detach().view(-1).cpu().numpy() if batch_indices is not None else loss.detach().item()
    return fval, gradf


def gen_candidates_scipy(initial_conditions: Tensor, acquisition_function, lower_bounds=None, upper_bounds=None,
                         inequality_constraints=None, equality_constraints=None, nonlinear_inequality_constraints=None,
                         options: dict | None = None, fixed_features=None, timeout_sec: float | None = None,
                         use_parallel_mode: bool | None = None) -> tuple[Tensor, Tensor]:
    if inequality_constraints or equality_constraints or nonlinear_inequality_constraints or fixed_features:
        raise UnsupportedError("botorch_b200.gen_candidates_scipy implements the box-bounded L-BFGS-B fast path only.")
    options = dict(options or {})
    options.setdefault("maxiter", 2000)
    if options.get("method", "L-BFGS-B") != "L-BFGS-B" or not options.get("with_grad", True):
        raise UnsupportedError("Only method='L-BFGS-B' with gradients is supported.")
    orig_shape = initial_conditions.shape
    if initial_conditions.ndim == 2:
        initial_conditions = initial_conditions.unsqueeze(0)
    clamped = columnwise_clamp(X=initial_conditions, lower=lower_bounds, upper=upper_bounds, raise_on_violation=True)

    def f(x):
        # optimiser rounds are a few hundred rows: the int8 mode runs them with its most accurate slice counts (FP64-level,
        # no variance-collapse re-routing; the choice depends on the PHASE, not on how a batch is chunked)
        with settings.int8_max_slices(True):
            return -acquisition_function(x)

    nb, q, d = clamped.shape
    x0 = _arrayify(clamped).reshape(nb, -1)
    lo = None if lower_bounds is None else torch.as_tensor(lower_bounds, dtype=torch.float64).expand(d).cpu().numpy()
    hi = None if upper_bounds is None else torch.as_tensor(upper_bounds, dtype=torch.float64).expand(d).cpu().numpy()
    bounds = None
    if lo is not None or hi is not None:
        lo_full = np.tile(lo if lo is not None else np.full(d, -np.inf), q)
        hi_full = np.tile(hi if hi is not None else np.full(d, np.inf), q)
        bounds = np.stack([lo_full, hi_full], axis=-1)
    minimize_opts = {k: v for k, v in options.items() if k in ("maxiter", "maxcor", "ftol", "pgtol", "maxls", "maxfun")}
    xs, fs, results = fmin_l_bfgs_b_batched(
        func=partial(_f_np_wrapper, f=f, shapeX=clamped.shape, device=initial_conditions.device,
                     dtype=initial_conditions.dtype),
        x0=x0, bounds=bounds, callback=options.get("callback"), pass_batch_indices=True, timeout_sec=timeout_sec,
        **minimize_opts)
    for res in results:
        _process_scipy_result(res=res, options=options)
    candidates = torch.from_numpy(xs).view_as(clamped).to(initial_conditions)
    clamped_candidates = columnwise_clamp(X=candidates, lower=lower_bounds, upper=upper_bounds,
                                          raise_on_violation=True).reshape(orig_shape)
    with torch.no_grad():
        batch_acquisition = acquisition_function(clamped_candidates)
    return clamped_candidates, batch_acquisition


def _process_scipy_result(res, options: dict) -> None:
    """Logs and warnings for one scipy result (reference generation/gen.py:711-746): running into the iteration /
    evaluation limit or the timeout is logged, every other failure raises an `OptimizationWarning`."""
    import logging

    logger = logging.getLogger("botorch")
    if "success" not in res.keys() or "status" not in res.keys():
        with warnings.catch_warnings():
            warnings.simplefilter("always", category=OptimizationWarning)
            warnings.warn("Optimization failed within `scipy.optimize.minimize` with no status returned to `res.`",
                          OptimizationWarning, stacklevel=3)
    elif not res.success:
        msg = res.message if isinstance(res.message, str) else res.message.decode("ascii")
        if "ITERATIONS REACHED LIMIT" in msg or "Iteration limit reached" in msg:
            logger.info(f"`scipy.optimize.minimize` exited by reaching the iteration limit of `maxiter: {options.get('maxiter')}`.")
        elif "EVALUATIONS EXCEEDS LIMIT" in msg:
            logger.info("`scipy.optimize.minimize` exited by reaching the function evaluation limit of "
                        f"`maxfun: {options.get('maxfun')}`.")
        elif "Optimization timed out after" in msg:
            logger.info(msg)
        else:
            with warnings.catch_warnings():
                warnings.simplefilter("always", category=OptimizationWarning)
                warnings.warn(f"Optimization failed within `scipy.optimize.minimize` with status {res.status} and "
                              f"message {msg}.", OptimizationWarning, stacklevel=3)
