"""MC objectives (reference: botorch/acquisition/objective.py:249-358).  `IdentityMCObjective` is fused
into the CUDA reduction; other objectives run on materialised samples."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Callable

import torch
from torch import Tensor
from torch.nn import Module


class PosteriorTransform(Module, ABC):
    scalarize: bool = False

    @abstractmethod
    def forward(self, posterior):
        ...


class MCAcquisitionObjective(Module, ABC):
    _is_mo: bool = False

    @abstractmethod
    def forward(self, samples: Tensor, X: Tensor | None = None) -> Tensor:
        """samples: sample_shape x batch_shape x q x m -> sample_shape x batch_shape x q."""

    def __call__(self, samples: Tensor, X: Tensor | None = None, *args, **kwargs) -> Tensor:
        return super().__call__(samples=samples, X=X, *args, **kwargs)


class IdentityMCObjective(MCAcquisitionObjective):
    def forward(self, samples: Tensor, X: Tensor | None = None) -> Tensor:
        return samples.squeeze(-1)


class LinearMCObjective(MCAcquisitionObjective):
    def __init__(self, weights: Tensor) -> None:
        super().__init__()
        if weights.dim() != 1:
            raise ValueError("weights must be a one-dimensional tensor.")
        self.register_buffer("weights", weights)

    def forward(self, samples: Tensor, X: Tensor | None = None) -> Tensor:
        if samples.shape[-1] != self.weights.shape[-1]:
            raise RuntimeError("Output shape of samples not equal to that of weights")
        return torch.einsum("...m, m", [samples, self.weights])


class GenericMCObjective(MCAcquisitionObjective):
    def __init__(self, objective: Callable[[Tensor, Tensor | None], Tensor]) -> None:
        super().__init__()
        self.objective = objective

    def forward(self, samples: Tensor, X: Tensor | None = None) -> Tensor:
        return self.objective(samples, X=X)
