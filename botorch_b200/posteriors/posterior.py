"""Posterior ABC (reference: botorch/posteriors/posterior.py:19-145)."""
from __future__ import annotations

from abc import ABC, abstractmethod

import torch
from torch import Tensor


class Posterior(ABC):
    def rsample_from_base_samples(self, sample_shape: torch.Size, base_samples: Tensor) -> Tensor:
        raise NotImplementedError(f"{self.__class__.__name__} does not implement `rsample_from_base_samples`.")

    @abstractmethod
    def rsample(self, sample_shape: torch.Size | None = None) -> Tensor:
        ...

    def sample(self, sample_shape: torch.Size | None = None) -> Tensor:
        with torch.no_grad():
            return self.rsample(sample_shape=sample_shape)

    @property
    @abstractmethod
    def device(self) -> torch.device:
        ...

    @property
    @abstractmethod
    def dtype(self) -> torch.dtype:
        ...

    @property
    def base_sample_shape(self) -> torch.Size:
        raise NotImplementedError

    @property
    def batch_range(self) -> tuple[int, int]:
        raise NotImplementedError

    def _extended_shape(self, sample_shape: torch.Size = torch.Size()) -> torch.Size:
        raise NotImplementedError
