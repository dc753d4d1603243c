"""`Normalize` input transform (reference: botorch/models/transforms/input.py:541-554, 629-770).
Affine, so the CUDA kernels fold it into the scaled inputs: u = ((x - offset) / coefficient) / lengthscale."""
from __future__ import annotations

import torch
from torch import Tensor
from torch.nn import Module


class Normalize(Module):
    def __init__(self, d: int, bounds: Tensor | None = None, min_range: float = 1e-8, learn_bounds: bool | None = None,
                 center: float = 0.5) -> None:
        super().__init__()
        self._d = d
        self.learn_coefficients = (bounds is None) if learn_bounds is None else learn_bounds
        self.min_range = min_range
        self.center = center
        if bounds is not None:
            offset = bounds[..., 0:1, :].to(torch.float64)
            coefficient = bounds[..., 1:2, :].to(torch.float64) - offset
        else:
            offset = torch.zeros(1, d, dtype=torch.float64)
            coefficient = torch.ones(1, d, dtype=torch.float64)
        self.register_buffer("_offset", offset + (0.5 - center) * coefficient)
        self.register_buffer("_coefficient", coefficient)

    @property
    def offset(self) -> Tensor:
        return self._offset

    @property
    def coefficient(self) -> Tensor:
        return self._coefficient

    @property
    def bounds(self) -> Tensor:
        return torch.cat([self.offset, self.offset + self.coefficient], dim=-2)

    def _update_coefficients(self, X: Tensor) -> None:
        lo = torch.amin(X, dim=tuple(range(X.ndim - 1))).unsqueeze(-2)
        rng = torch.amax(X, dim=tuple(range(X.ndim - 1))).unsqueeze(-2) - lo
        tiny = rng < self.min_range
        self._coefficient = torch.where(tiny, 1.0, rng).to(self._coefficient)
        self._offset = (torch.where(tiny, 0.0, lo) + (0.5 - self.center) * self._coefficient).to(self._offset)

    def transform(self, X: Tensor) -> Tensor:
        if self.learn_coefficients and self.training:
            self._update_coefficients(X)
        return (X - self.offset.to(X)) / self.coefficient.to(X)

    forward = transform

    def untransform(self, X: Tensor) -> Tensor:
        return X * self.coefficient.to(X) + self.offset.to(X)
