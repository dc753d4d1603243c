"""Synthetic objectives used to generate the BASELINE configurations (formulas as in
botorch/test_functions/synthetic.py:137-181 Ackley, :387-487 Hartmann).  Data generators only."""
from __future__ import annotations

import math

import torch
from torch import Tensor

# Hartmann-6 constants (standard literature values)
_H6_ALPHA = (1.0, 1.2, 3.0, 3.2)
_H6_A = ((10, 3, 17, 3.5, 1.7, 8), (0.05, 10, 17, 0.1, 8, 14), (3, 3.5, 1.7, 10, 17, 8), (17, 8, 0.05, 10, 0.1, 14))
_H6_P = ((1312, 1696, 5569, 124, 8283, 5886), (2329, 4135, 8307, 3736, 1004, 9991),
         (2348, 1451, 3522, 2883, 3047, 6650), (4047, 8828, 8732, 5743, 1091, 381))


class _Synthetic:
    def __init__(self, dim: int, bounds, noise_std: float | None, negate: bool, dtype=torch.double) -> None:
        self.dim, self.noise_std, self.negate = dim, noise_std, negate
        self.bounds = torch.tensor(bounds, dtype=dtype).t().contiguous()  # 2 x d

    def evaluate_true(self, X: Tensor) -> Tensor:
        raise NotImplementedError

    def __call__(self, X: Tensor, noise: bool = True) -> Tensor:
        f = self.evaluate_true(X)
        if noise and self.noise_std is not None:
            f = f + self.noise_std * torch.randn_like(f)
        return -f if self.negate else f


class Ackley(_Synthetic):
    """f(x) = -20 exp(-0.2 sqrt(mean x^2)) - exp(mean cos(2 pi x)) + 20 + e on [-32.768, 32.768]^d."""

    def __init__(self, dim: int = 2, noise_std: float | None = None, negate: bool = False, bounds=None) -> None:
        super().__init__(dim, bounds or [(-32.768, 32.768)] * dim, noise_std, negate)

    def evaluate_true(self, X: Tensor) -> Tensor:
        rms_term = -20.0 * torch.exp(-0.2 * torch.linalg.norm(X, dim=-1) / math.sqrt(self.dim))
        cos_term = -torch.exp(torch.cos(2.0 * math.pi * X).mean(dim=-1))
        return rms_term + cos_term + 20.0 + math.e


class Hartmann(_Synthetic):
    """Six-dimensional Hartmann function on [0, 1]^6, global minimum -3.32237."""

    def __init__(self, dim: int = 6, noise_std: float | None = None, negate: bool = False, bounds=None) -> None:
        if dim != 6:
            raise ValueError("botorch_b200 ships the 6-dimensional Hartmann function only.")
        super().__init__(dim, bounds or [(0.0, 1.0)] * dim, noise_std, negate)
        self.ALPHA = torch.tensor(_H6_ALPHA, dtype=torch.double)
        self.A = torch.tensor(_H6_A, dtype=torch.double)
        self.P = torch.tensor(_H6_P, dtype=torch.double) * 1e-4

    def evaluate_true(self, X: Tensor) -> Tensor:
        diff = X.unsqueeze(-2) - self.P.to(X)
        inner = (self.A.to(X) * diff.square()).sum(dim=-1)
        return -(self.ALPHA.to(X) * torch.exp(-inner)).sum(dim=-1)
