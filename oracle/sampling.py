"""Base-sample construction restated (TEST INFRASTRUCTURE ONLY).

Follows the reference:
  botorch/sampling/qmc.py:59-97           NormalQMCEngine.draw (inverse transform branch)
  botorch/utils/sampling.py:74-111        draw_sobol_samples
  botorch/utils/sampling.py:114-143       draw_sobol_normal_samples
  botorch/sampling/normal.py:182-213      SobolQMCNormalSampler._construct_base_samples
  botorch/sampling/normal.py:68-135       NormalMCSampler._update_base_samples (first r columns frozen)
  botorch/sampling/base.py:98-116         collapsed t-batch shape  S x 1 x q (x m)
`torch.quasirandom.SobolEngine` is part of torch, so draws are bit-identical to the reference's.
"""
from __future__ import annotations

import math

import torch
from torch import Tensor
from torch.quasirandom import SobolEngine


def draw_sobol_normal_samples(d: int, n: int, dtype=torch.float64, seed: int | None = None) -> Tensor:
    engine = SobolEngine(dimension=d, scramble=True, seed=seed)
    samples = engine.draw(n, dtype=dtype)
    v = 0.5 + (1 - torch.finfo(samples.dtype).eps) * (samples - 0.5)
    return torch.erfinv(2 * v - 1) * math.sqrt(2)


def draw_sobol_samples(bounds: Tensor, n: int, q: int, seed: int | None = None) -> Tensor:
    d = bounds.shape[-1]
    engine = SobolEngine(q * d, scramble=True, seed=seed)
    raw = engine.draw(n, dtype=bounds.dtype).view(n, q, d)
    return raw * (bounds[1] - bounds[0]) + bounds[0]  # utils/transforms.py:99-129 unnormalize


def qlogei_base_samples(S: int, q: int, seed: int, dtype=torch.float64) -> Tensor:
    """S x 1 x q x 1 base samples for a single-output posterior over q points."""
    return draw_sobol_normal_samples(d=q, n=S, dtype=dtype, seed=seed).view(S, 1, q, 1)


def qlognei_base_samples(S: int, r: int, q: int, seed: int, dtype=torch.float64) -> Tensor:
    """S x 1 x (r+q) x 1: fresh (r+q)-dim draw whose first r columns are overwritten with the r-dim
    draw used when the baseline was sampled (same seed) -- sampling/normal.py:79-135."""
    base = draw_sobol_normal_samples(d=r, n=S, dtype=dtype, seed=seed).view(S, 1, r, 1)
    full = draw_sobol_normal_samples(d=r + q, n=S, dtype=dtype, seed=seed).view(S, 1, r + q, 1).clone()
    full[..., :r, :] = base
    return full
