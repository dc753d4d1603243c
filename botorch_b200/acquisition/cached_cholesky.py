"""Cached-root sampling mixin (reference: botorch/acquisition/cached_cholesky.py:34-192)."""
from __future__ import annotations

import warnings

import torch
from torch import Tensor

from ..exceptions.errors import NanError, NotPSDError
from ..exceptions.warnings import BotorchWarning
from ..models.gp_regression import SingleTaskGP
from ..sampling.base import MCSampler
from ..utils.low_rank import sample_cached_cholesky
from .acquisition import MCSamplerMixin


def supports_cache_root(model) -> bool:
    """Exact single-task GPs with linear outcome transforms support the cached root (:34-50)."""
    return isinstance(model, SingleTaskGP)


class CachedCholeskyMCSamplerMixin(MCSamplerMixin):
    def __init__(self, model, cache_root: bool | None = None, sampler: MCSampler | None = None) -> None:
        MCSamplerMixin.__init__(self, sampler=sampler)
        if cache_root is None:
            cache_root = supports_cache_root(model)
        elif cache_root and not supports_cache_root(model):
            warnings.warn(f"`cache_root` is only supported for exact GPs; got {type(model).__name__}.", RuntimeWarning,
                          stacklevel=3)
            cache_root = False
        self._cache_root = cache_root

    def _compute_root_decomposition(self, posterior) -> Tensor:
        return posterior.distribution.scale_tril

    def _get_f_X_samples(self, posterior, q_in: int) -> Tensor:
        if self._cache_root and hasattr(self, "_baseline_L"):
            try:
                return sample_cached_cholesky(posterior=posterior, baseline_L=self._baseline_L, q=q_in,
                                              base_samples=self.sampler.base_samples,
                                              sample_shape=self.sampler.sample_shape)
            except (NanError, NotPSDError):
                warnings.warn("Low-rank cholesky updates failed due NaNs or due to an ill-conditioned covariance "
                              "matrix. Falling back to standard sampling.", BotorchWarning, stacklevel=3)
        samples = self.get_posterior_samples(posterior)
        return samples[..., -q_in:, :]

    def _set_sampler(self, q_in: int, posterior) -> None:
        if self.q_in != q_in and self.base_sampler is not None:
            self.sampler._update_base_samples(posterior=posterior, base_sampler=self.base_sampler)
            self.q_in = q_in
