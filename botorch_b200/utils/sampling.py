"""Quasi-random draws with the reference's seed semantics (botorch/utils/sampling.py).

`torch.quasirandom.SobolEngine` (torch's own generator) does the scrambling, so a given seed yields the reference's
points bit for bit, which is what makes selected restart indices comparable.  The base samples of the MC samplers are
drawn on the host once per acquisition function; the `raw_samples x q x d` Sobol cloud of `gen_batch_initial_conditions`
is generated on the device from the engine's state (`mcacq_sobol_draw`) when the bounds live there.
"""
from __future__ import annotations

import math
from contextlib import contextmanager

import torch
from torch import Tensor
from torch.quasirandom import SobolEngine


@contextmanager
def manual_seed(seed: int | None = None):
    """Temporarily seed torch's global CPU generator (reference :51-71)."""
    old_state = torch.random.get_rng_state()
    try:
        if seed is not None:
            torch.random.manual_seed(seed)
        yield
    finally:
        if seed is not None:
            torch.random.set_rng_state(old_state)


def _device_sobol(engine: SobolEngine, n: int, device, dtype) -> Tensor:
    """`engine.draw(n)` of a FRESH engine, generated on `device` by `mcacq_sobol_draw` from the engine's own scrambled
    direction numbers and shift.  Bit-identical to the host draw: integer XORs scaled by 2^-30; row 0 is the engine's
    `_first_point`, which torch rounds through its default dtype."""
    from .. import _lib

    if engine.num_generated != 0:
        raise ValueError("the device draw reproduces a fresh engine only")
    out = torch.empty(n, engine.dimension, device=device, dtype=torch.float64)
    if n == 0:
        return out.to(dtype)
    ss = engine.sobolstate.to(device=device, dtype=torch.int64).contiguous()
    shift = engine.shift.to(device=device, dtype=torch.int64).contiguous()
    with torch.cuda.device(device):
        _lib.check(_lib.lib().mcacq_sobol_draw(ss.data_ptr(), shift.data_ptr(), engine.dimension, n, 0, out.data_ptr(),
                                               _lib.stream_ptr()), "mcacq_sobol_draw")
    out[0] = engine._first_point.to(dtype).to(device=device, dtype=torch.float64).reshape(-1)
    return out.to(dtype)


_ENGINES: dict[tuple[int, int], SobolEngine] = {}


def _fresh_engine(dimension: int, seed: int | None) -> SobolEngine:
    """A scrambled engine that has not drawn yet.  For an explicit seed the (never advanced) engine is kept: torch spends
    10-50 ms of host time per construction on the scrambling matrices, and the device draw only reads its state."""
    if seed is None:
        return SobolEngine(dimension, scramble=True, seed=None)
    key = (dimension, int(seed))
    eng = _ENGINES.get(key)
    if eng is None:
        if len(_ENGINES) >= 16:
            _ENGINES.pop(next(iter(_ENGINES)))
        eng = _ENGINES[key] = SobolEngine(dimension, scramble=True, seed=seed)
    return eng


def draw_sobol_samples(bounds: Tensor, n: int, q: int, batch_shape=None, seed: int | None = None) -> Tensor:
    """`n x batch_shape x q x d` scrambled-Sobol points inside `bounds` (reference :74-111).  With `bounds` on a CUDA device
    the points are generated there (same engine, same scrambling, bit-identical values) instead of drawn on the host and
    copied."""
    batch_shape = torch.Size(batch_shape or ())
    nb = batch_shape.numel()
    d = bounds.shape[-1]
    if bounds.is_cuda and bounds.dtype in (torch.float64, torch.float32):
        raw = _device_sobol(_fresh_engine(q * d, seed), nb * n, bounds.device, bounds.dtype)
    else:
        raw = SobolEngine(q * d, scramble=True, seed=seed).draw(nb * n, dtype=bounds.dtype)
    raw = raw.view(*batch_shape, n, q, d).to(device=bounds.device)
    if len(batch_shape) > 0:
        raw = raw.permute(-3, *range(len(batch_shape)), -2, -1)
    return raw * (bounds[1] - bounds[0]) + bounds[0]


def draw_sobol_normal_samples(d: int, n: int, device=None, dtype=None, seed: int | None = None) -> Tensor:
    """`n x d` qMC N(0, I) draws by inverse-CDF of scrambled Sobol points (reference :114-143 and
    sampling/qmc.py:77-83: v = 0.5 + (1 - eps)(u - 0.5); z = sqrt(2) erfinv(2v - 1))."""
    dtype = torch.get_default_dtype() if dtype is None else dtype
    u = SobolEngine(dimension=d, scramble=True, seed=seed).draw(n, dtype=dtype)
    v = 0.5 + (1 - torch.finfo(u.dtype).eps) * (u - 0.5)
    return (torch.erfinv(2 * v - 1) * math.sqrt(2)).to(device=device)


def batched_multinomial(weights: Tensor, num_samples: int, replacement: bool = False, generator=None) -> Tensor:
    """`torch.multinomial` over the last dim for arbitrarily batched weights (reference :317-351)."""
    flat = weights.reshape(-1, weights.shape[-1])
    out = torch.multinomial(flat, num_samples=num_samples, replacement=replacement, generator=generator)
    return out.view(*weights.shape[:-1], num_samples)


def boltzmann_sample(function_values: Tensor, num_samples: int, eta: float, replacement: bool = False,
                     temp_decrease: float = 0.5) -> Tensor:
    """Indices drawn with probability proportional to exp(eta * zscore(f)) (reference :1084-1115)."""
    from .transforms import standardize

    norm_weights = standardize(function_values)
    weights = torch.exp(eta * norm_weights)
    while torch.isinf(weights).any():
        eta *= temp_decrease
        weights = torch.exp(eta * norm_weights)
    return batched_multinomial(weights=weights, num_samples=num_samples, replacement=replacement)


def sample_truncated_normal_perturbations(X: Tensor, n_discrete_points: int, sigma: float, bounds: Tensor,
                                          qmc: bool = True) -> Tensor:
    """`n_discrete_points` points N(x, sigma^2 I) around randomly chosen rows x of `X`, truncated to the box by inverse-CDF
    sampling in the normalised cube (reference utils/sampling.py:1118-1166)."""
    Xn = (X - bounds[0]) / (bounds[1] - bounds[0])
    d = Xn.shape[1]
    if Xn.shape[0] > 1:
        Xn = Xn[torch.randint(Xn.shape[0], (n_discrete_points,), device=Xn.device)]
    if qmc:
        std_bounds = torch.zeros(2, d, dtype=Xn.dtype, device=Xn.device)
        std_bounds[1] = 1
        u = draw_sobol_samples(bounds=std_bounds, n=n_discrete_points, q=1).squeeze(1)
    else:
        u = torch.rand((n_discrete_points, d), dtype=Xn.dtype, device=Xn.device)
    normal = torch.distributions.Normal(0, 1)
    cdf_alpha = normal.cdf(-Xn / sigma)
    perturbation = normal.icdf(cdf_alpha + u * (normal.cdf((1 - Xn) / sigma) - cdf_alpha)) * sigma
    return (Xn + perturbation).clamp(0.0, 1.0) * (bounds[1] - bounds[0]) + bounds[0]


def sample_perturbed_subset_dims(X: Tensor, bounds: Tensor, n_discrete_points: int, sigma: float = 1e-1, qmc: bool = True,
                                 prob_perturb: float | None = None) -> Tensor:
    """Perturb a random subset of the dimensions (probability min(20/d, 1) each, at least `ceil(d p)` when none was hit)
    of randomly chosen rows of `X` (reference utils/sampling.py:1169-1242)."""
    from ..exceptions.errors import BotorchTensorDimensionError

    if bounds.ndim != 2:
        raise BotorchTensorDimensionError("bounds must be a `2 x d`-dim tensor.")
    if X.ndim != 2:
        raise BotorchTensorDimensionError("X must be a `n x d`-dim tensor.")
    d = bounds.shape[-1]
    if prob_perturb is None:
        prob_perturb = min(20.0 / d, 1.0)
    if X.shape[0] == 1:
        X_cand = X.repeat(n_discrete_points, 1)
    else:
        X_cand = X[torch.randint(X.shape[0], (n_discrete_points,), device=X.device)]
    pert = sample_truncated_normal_perturbations(X=X_cand, n_discrete_points=n_discrete_points, sigma=sigma, bounds=bounds,
                                                 qmc=qmc)
    mask = torch.rand(n_discrete_points, d, dtype=bounds.dtype, device=bounds.device) <= prob_perturb
    ind = (~mask).all(dim=-1).nonzero()
    n_perturb = math.ceil(d * prob_perturb)
    perturb_mask = torch.zeros(d, dtype=mask.dtype, device=mask.device)
    perturb_mask[:n_perturb].fill_(1)
    for idx in ind:
        mask[idx] = perturb_mask[torch.randperm(d, device=bounds.device)]
    X_cand[mask] = pert[mask]
    return X_cand
