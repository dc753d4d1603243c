"""Multivariate-normal posterior over q points (reference: botorch/posteriors/gpytorch.py:35-180).

`MultivariateNormal` is the minimal stand-in for gpytorch's distribution class: a dense mean / covariance
pair on the device with a lazily cached `psd_safe_cholesky` root.  Sampling follows
`MultivariateNormal.rsample(base_samples=...)`: `y = mean + chol(covar) z`.
"""
from __future__ import annotations

import torch
from torch import Tensor

from ..models.prediction_strategy import psd_safe_cholesky
from .posterior import Posterior

MIN_VARIANCE = 1e-10  # gpytorch settings.min_variance (fp64)


class MultivariateNormal:
    def __init__(self, mean: Tensor, covariance_matrix: Tensor) -> None:
        self.loc = mean
        self._covar = covariance_matrix
        self._root = None

    @property
    def mean(self) -> Tensor:
        return self.loc

    @property
    def covariance_matrix(self) -> Tensor:
        return self._covar

    lazy_covariance_matrix = covariance_matrix

    @property
    def batch_shape(self) -> torch.Size:
        return self.loc.shape[:-1]

    @property
    def event_shape(self) -> torch.Size:
        return self.loc.shape[-1:]

    base_sample_shape = event_shape

    @property
    def variance(self) -> Tensor:
        return self._covar.diagonal(dim1=-1, dim2=-2).clamp_min(MIN_VARIANCE)

    @property
    def scale_tril(self) -> Tensor:
        if self._root is None:
            self._root = psd_safe_cholesky(self._covar, max_tries=6)
        return self._root

    def rsample(self, sample_shape: torch.Size = torch.Size(), base_samples: Tensor | None = None) -> Tensor:
        root = self.scale_tril
        if base_samples is None:
            z = torch.randn(*sample_shape, *self.loc.shape, dtype=self.loc.dtype, device=self.loc.device)
        else:
            z = base_samples
        z = z.reshape(-1, *self.loc.shape[:-1], root.shape[-1])
        z = z.permute(*range(1, self.loc.dim() + 1), 0)
        res = root.matmul(z) + self.loc.unsqueeze(-1)
        res = res.permute(-1, *range(self.loc.dim())).contiguous()
        return res.view(torch.Size(sample_shape) + self.loc.shape)


class GPyTorchPosterior(Posterior):
    def __init__(self, distribution: MultivariateNormal) -> None:
        self.distribution = distribution
        self._is_mt = False

    @property
    def mvn(self) -> MultivariateNormal:
        return self.distribution

    @property
    def device(self) -> torch.device:
        return self.distribution.loc.device

    @property
    def dtype(self) -> torch.dtype:
        return self.distribution.loc.dtype

    @property
    def batch_shape(self) -> torch.Size:
        return self.distribution.batch_shape

    @property
    def base_sample_shape(self) -> torch.Size:
        return self.distribution.batch_shape + self.distribution.base_sample_shape

    @property
    def batch_range(self) -> tuple[int, int]:
        return (0, -1)

    def _extended_shape(self, sample_shape: torch.Size = torch.Size()) -> torch.Size:
        return sample_shape + self.distribution.batch_shape + self.distribution.event_shape + torch.Size([1])

    def rsample_from_base_samples(self, sample_shape: torch.Size, base_samples: Tensor) -> Tensor:
        if base_samples.shape[: len(sample_shape)] != sample_shape:
            raise RuntimeError(
                f"`sample_shape` disagrees with shape of `base_samples`. Got {sample_shape=} and {base_samples.shape=}.")
        return self.distribution.rsample(sample_shape=sample_shape, base_samples=base_samples).unsqueeze(-1)

    def rsample(self, sample_shape: torch.Size | None = None) -> Tensor:
        if sample_shape is None:
            sample_shape = torch.Size([1])
        return self.distribution.rsample(sample_shape=sample_shape).unsqueeze(-1)

    @property
    def mean(self) -> Tensor:
        return self.distribution.mean.unsqueeze(-1)

    @property
    def variance(self) -> Tensor:
        return self.distribution.variance.unsqueeze(-1)

    @property
    def covariance_matrix(self) -> Tensor:
        return self.distribution.covariance_matrix
