"""Model base class (reference: botorch/models/model.py:79-250): `posterior`, `batch_shape`,
`num_outputs`, `transform_inputs`."""
from __future__ import annotations

from abc import ABC

import torch
from torch import Tensor
from torch.nn import Module


class Model(Module, ABC):
    _is_fully_bayesian = False
    _is_ensemble = False

    def posterior(self, X: Tensor, output_indices=None, observation_noise=False, posterior_transform=None):
        raise NotImplementedError(f"{self.__class__.__name__} does not implement a `posterior` method.")

    @property
    def batch_shape(self) -> torch.Size:
        raise NotImplementedError

    @property
    def num_outputs(self) -> int:
        raise NotImplementedError

    def transform_inputs(self, X: Tensor, input_transform: Module | None = None) -> Tensor:
        if input_transform is not None:
            input_transform.to(X)
            return input_transform(X)
        tf = getattr(self, "input_transform", None)
        return X if tf is None else tf(X)
