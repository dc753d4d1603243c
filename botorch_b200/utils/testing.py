"""Deterministic stand-ins for a model and its posterior: what the reference's acquisition tests are written against
(`MockPosterior` / `MockModel`, botorch/utils/testing.py:477-665).  A `MockPosterior` hands back the samples it was built
with, whatever base samples it is given, so the value of an MC acquisition function is known in closed form; the tests in
`tests/test_reference_known_answers.py` restate the reference's expectations (test/acquisition/test_logei.py:100-455,
test_monte_carlo.py) against this package's classes."""
from __future__ import annotations

import torch
from torch import Tensor

from ..models.model import Model
from ..posteriors.posterior import Posterior


class MockPosterior(Posterior):
    def __init__(self, mean: Tensor | None = None, variance: Tensor | None = None, samples: Tensor | None = None,
                 base_shape: torch.Size | None = None, batch_range: tuple[int, int] | None = None) -> None:
        self._mean, self._variance, self._samples = mean, variance, samples
        self._base_shape = base_shape
        self._batch_range = batch_range or (0, -2)

    def _first(self) -> Tensor | None:
        return next((t for t in (self._mean, self._variance, self._samples) if torch.is_tensor(t)), None)

    @property
    def device(self) -> torch.device:
        t = self._first()
        return torch.device("cpu") if t is None else t.device

    @property
    def dtype(self) -> torch.dtype:
        t = self._first()
        return torch.float32 if t is None else t.dtype

    @property
    def batch_shape(self) -> torch.Size:
        t = self._first()
        if t is None:
            raise NotImplementedError
        return t.shape[:-2]

    @property
    def base_sample_shape(self) -> torch.Size:
        if self._base_shape is not None:
            return self._base_shape
        t = next((t for t in (self._samples, self._mean, self._variance) if t is not None), None)
        return torch.Size() if t is None else t.shape

    @property
    def batch_range(self) -> tuple[int, int]:
        return self._batch_range

    def _extended_shape(self, sample_shape: torch.Size = torch.Size()) -> torch.Size:
        return sample_shape + self.base_sample_shape

    @property
    def mean(self):
        return self._mean

    @property
    def variance(self):
        return self._variance

    def rsample(self, sample_shape: torch.Size | None = None) -> Tensor:
        return self._samples.expand(self._extended_shape(torch.Size() if sample_shape is None else sample_shape))

    def rsample_from_base_samples(self, sample_shape: torch.Size, base_samples: Tensor) -> Tensor:
        if base_samples.shape[: len(sample_shape)] != sample_shape:
            raise RuntimeError("`sample_shape` disagrees with shape of `base_samples`. "
                               f"Got {sample_shape=} and {base_samples.shape=}.")
        return self.rsample(sample_shape)


class MockModel(Model):
    """`posterior(X)` ignores X and returns the posterior it was built with (transformed if a transform is given)."""

    def __init__(self, posterior: MockPosterior) -> None:
        super().__init__()
        self._posterior = posterior

    def posterior(self, X: Tensor, output_indices=None, observation_noise=False, posterior_transform=None):
        return self._posterior if posterior_transform is None else posterior_transform(self._posterior)

    @property
    def num_outputs(self) -> int:
        shape = self._posterior._extended_shape()
        return shape[-1] if len(shape) > 0 else 0

    @property
    def batch_shape(self) -> torch.Size:
        return self._posterior._extended_shape()[:-2]
