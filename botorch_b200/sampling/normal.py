"""Normal base-sample samplers (reference: botorch/sampling/normal.py:34-213).

`SobolQMCNormalSampler` draws `S x 1 x q (x m)` scrambled-Sobol normal base samples once per shape
(:182-213) and re-uses them for every t-batch; `_update_base_samples` (:68-135) freezes the leading
columns to those of a baseline sampler (qLogNEI's cached-root path).
"""
from __future__ import annotations

import torch
from torch import Tensor
from torch.quasirandom import SobolEngine

from ..exceptions.errors import UnsupportedError
from ..utils.sampling import draw_sobol_normal_samples, manual_seed
from .base import MCSampler


class NormalMCSampler(MCSampler):
    def forward(self, posterior) -> Tensor:
        self._construct_base_samples(posterior=posterior)
        return posterior.rsample_from_base_samples(
            sample_shape=self.sample_shape,
            base_samples=self.base_samples.expand(self._get_extended_base_sample_shape(posterior=posterior)),
        )

    def _construct_base_samples(self, posterior) -> None:  # pragma: no cover
        raise NotImplementedError

    def _update_base_samples(self, posterior, base_sampler: MCSampler) -> None:
        self._instance_check(base_sampler=base_sampler)
        self._construct_base_samples(posterior=posterior)
        if base_sampler.base_samples is None:
            return
        cur = base_sampler.base_samples.detach().clone()
        base_ndims = cur.dim() - 1
        target = self._get_collapsed_shape(posterior=posterior)
        view_shape = self.sample_shape + torch.Size([1] * (len(target) - cur.dim())) + cur.shape[-base_ndims:]
        expanded = cur.view(view_shape).expand(target[:-base_ndims] + cur.shape[-base_ndims:])
        single_output = (len(posterior.base_sample_shape) - len(posterior.batch_shape)) == 1
        if single_output:
            self.base_samples[..., : cur.shape[-1]] = expanded
        else:
            self.base_samples[..., : cur.shape[-2], :] = expanded


class IIDNormalSampler(NormalMCSampler):
    """iid N(0, 1) base samples under `manual_seed(self.seed)` (reference :138-170)."""

    def _construct_base_samples(self, posterior) -> None:
        target = self._get_collapsed_shape(posterior=posterior)
        if self.base_samples is None or self.base_samples.shape != target:
            with manual_seed(seed=self.seed):
                base = torch.randn(target, device=posterior.device, dtype=posterior.dtype)
            self.register_buffer("base_samples", base)
        if self.base_samples.device != posterior.device or self.base_samples.dtype != posterior.dtype:
            self.to(device=posterior.device, dtype=posterior.dtype)


class SobolQMCNormalSampler(NormalMCSampler):
    """Scrambled-Sobol qMC normal base samples (reference :173-213)."""

    def _construct_base_samples(self, posterior) -> None:
        target = self._get_collapsed_shape(posterior=posterior)
        if self.base_samples is None or self.base_samples.shape != target:
            out_dim = target[len(self.sample_shape):].numel()
            if out_dim > SobolEngine.MAXDIM:
                raise UnsupportedError(
                    f"SobolQMCSampler only supports dimensions `q * o <= {SobolEngine.MAXDIM}`. Requested: {out_dim}")
            base = draw_sobol_normal_samples(d=out_dim, n=self.sample_shape.numel(), device=posterior.device,
                                             dtype=posterior.dtype, seed=self.seed)
            self.register_buffer("base_samples", base.view(target))
        self.to(device=posterior.device, dtype=posterior.dtype)
