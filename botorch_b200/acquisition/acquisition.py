"""Acquisition base classes (reference: botorch/acquisition/acquisition.py:24-146)."""
from __future__ import annotations

import warnings
from abc import ABC, abstractmethod

import torch
from torch import Tensor
from torch.nn import Module

from ..exceptions.warnings import BotorchWarning
from ..sampling.base import MCSampler
from ..sampling.get_sampler import get_sampler


class AcquisitionFunction(Module, ABC):
    """`forward(X: (b) x q x d) -> (b)` -- the drop-in boundary."""

    _log: bool = False

    def __init__(self, model) -> None:
        super().__init__()
        self.model = model

    def set_X_pending(self, X_pending: Tensor | None = None) -> None:
        if X_pending is not None:
            if X_pending.requires_grad:
                warnings.warn("Pending points require a gradient but the acquisition function will not provide a "
                              "gradient to these points.", BotorchWarning, stacklevel=2)
            self.X_pending = X_pending.detach().clone()
        else:
            self.X_pending = X_pending

    @abstractmethod
    def forward(self, X: Tensor) -> Tensor:
        ...


class MCSamplerMixin(ABC):
    """Lazily creates a 512-sample sampler on first use (reference :109-146)."""

    _default_sample_shape = torch.Size([512])

    def __init__(self, sampler: MCSampler | None = None) -> None:
        self.sampler = sampler

    def get_posterior_samples(self, posterior) -> Tensor:
        if self.sampler is None:
            self.sampler = get_sampler(posterior=posterior, sample_shape=self._default_sample_shape)
        return self.sampler(posterior=posterior)

    @property
    def sample_shape(self) -> torch.Size:
        return self.sampler.sample_shape if self.sampler is not None else self._default_sample_shape
