"""`get_sampler` (reference: botorch/sampling/get_sampler.py:44-88): Sobol for normal posteriors whose
collapsed base-sample dimension fits the Sobol engine, iid otherwise."""
from __future__ import annotations

import torch
from torch.quasirandom import SobolEngine

from .base import MCSampler
from .normal import IIDNormalSampler, SobolQMCNormalSampler


def get_sampler(posterior, sample_shape: torch.Size, *, seed: int | None = None) -> MCSampler:
    sampler = SobolQMCNormalSampler(sample_shape=sample_shape, seed=seed)
    collapsed = sampler._get_collapsed_shape(posterior=posterior)
    if collapsed[len(sample_shape):].numel() > SobolEngine.MAXDIM:
        sampler = IIDNormalSampler(sample_shape=sample_shape, seed=seed)
    return sampler
