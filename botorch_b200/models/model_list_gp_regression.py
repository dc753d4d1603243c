"""`ModelListGP`: m independent single-output GPs (reference: botorch/models/model_list_gp_regression.py:22-138,
botorch/models/gpytorch.py:786-884).  The joint posterior has a block-diagonal covariance (one q x q block per
output), which is what `MultitaskMultivariateNormal.from_independent_mvns` builds in the reference; here it is kept
as the list of per-output blocks, each produced by the CUDA posterior kernel."""
from __future__ import annotations

import torch
from torch import Tensor
from torch.nn import ModuleList

from ..posteriors.gpytorch import MultivariateNormal
from ..posteriors.posterior import Posterior
from .model import Model


class IndependentOutputsPosterior(Posterior):
    """Posterior over (q points x m outputs) with independent outputs; base samples have shape `... x q x m`
    (the non-interleaved layout of the reference's multi-task MVN: output k uses base-sample column k)."""

    def __init__(self, mvns: list[MultivariateNormal]) -> None:
        self.mvns = mvns
        self._is_mt = True

    @property
    def device(self) -> torch.device:
        return self.mvns[0].loc.device

    @property
    def dtype(self) -> torch.dtype:
        return self.mvns[0].loc.dtype

    @property
    def batch_shape(self) -> torch.Size:
        return self.mvns[0].batch_shape

    @property
    def base_sample_shape(self) -> torch.Size:
        return self.mvns[0].batch_shape + self.mvns[0].event_shape + torch.Size([len(self.mvns)])

    @property
    def batch_range(self) -> tuple[int, int]:
        return (0, -2)

    def _extended_shape(self, sample_shape: torch.Size = torch.Size()) -> torch.Size:
        return sample_shape + self.base_sample_shape

    @property
    def mean(self) -> Tensor:
        return torch.stack([m.mean for m in self.mvns], dim=-1)

    @property
    def variance(self) -> Tensor:
        return torch.stack([m.variance for m in self.mvns], dim=-1)

    def rsample_from_base_samples(self, sample_shape: torch.Size, base_samples: Tensor) -> Tensor:
        if base_samples.shape[: len(sample_shape)] != sample_shape:
            raise RuntimeError(f"`sample_shape` disagrees with shape of `base_samples`. Got {sample_shape=} and "
                               f"{base_samples.shape=}.")
        outs = [m.rsample(sample_shape=sample_shape, base_samples=base_samples[..., k].contiguous())
                for k, m in enumerate(self.mvns)]
        return torch.stack(outs, dim=-1)

    def rsample(self, sample_shape: torch.Size | None = None) -> Tensor:
        sample_shape = torch.Size([1]) if sample_shape is None else sample_shape
        return torch.stack([m.rsample(sample_shape=sample_shape) for m in self.mvns], dim=-1)


class ModelListGP(Model):
    def __init__(self, *gp_models) -> None:
        super().__init__()
        self.models = ModuleList(gp_models)

    @property
    def num_outputs(self) -> int:
        return sum(m.num_outputs for m in self.models)

    @property
    def batch_shape(self) -> torch.Size:
        return torch.Size()

    def posterior(self, X: Tensor, output_indices=None, observation_noise=False, posterior_transform=None):
        idx = range(len(self.models)) if output_indices is None else output_indices
        mvns = [self.models[i].posterior(X, observation_noise=observation_noise).distribution for i in idx]
        posterior = IndependentOutputsPosterior(mvns)
        return posterior if posterior_transform is None else posterior_transform(posterior)
