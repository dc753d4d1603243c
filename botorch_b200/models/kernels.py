"""Hyper-parameter holders with gpytorch's attribute names (`lengthscale`, `outputscale`, `noise`,
`constant`).  Fitting is out of scope: these modules only carry the values the
CUDA kernels consume.  Defaults follow botorch/models/utils/gpytorch_modules.py:74-133."""
from __future__ import annotations

import math

import torch
from torch import Tensor
from torch.nn import Module, Parameter

MIN_INFERRED_NOISE_LEVEL = 1e-4  # gpytorch_modules.py:29


class Kernel(Module):
    kernel_id: int = -1

    def __init__(self, ard_num_dims: int | None = None, lengthscale: Tensor | float | None = None) -> None:
        super().__init__()
        self.ard_num_dims = ard_num_dims
        d = 1 if ard_num_dims is None else ard_num_dims
        ls = torch.ones(1, d, dtype=torch.float64) if lengthscale is None else torch.as_tensor(
            lengthscale, dtype=torch.float64).reshape(1, -1)
        self.raw_lengthscale = Parameter(ls.clone())

    @property
    def lengthscale(self) -> Tensor:
        return self.raw_lengthscale

    @lengthscale.setter
    def lengthscale(self, value) -> None:
        with torch.no_grad():
            self.raw_lengthscale.copy_(torch.as_tensor(value, dtype=self.raw_lengthscale.dtype).reshape(1, -1))


class RBFKernel(Kernel):
    """k(x, x') = exp(-0.5 * ||(x - x') / l||^2)."""

    kernel_id = 0


class MaternKernel(Kernel):
    """Matern nu = 5/2: (1 + sqrt5 r + 5/3 r^2) exp(-sqrt5 r), r = ||(x - x') / l||."""

    kernel_id = 1

    def __init__(self, nu: float = 2.5, ard_num_dims: int | None = None, lengthscale=None) -> None:
        if nu != 2.5:
            raise NotImplementedError("botorch_b200 implements the Matern kernel for nu=2.5 only.")
        super().__init__(ard_num_dims=ard_num_dims, lengthscale=lengthscale)
        self.nu = nu


class ScaleKernel(Module):
    def __init__(self, base_kernel: Kernel, outputscale: float = 1.0) -> None:
        super().__init__()
        self.base_kernel = base_kernel
        self.raw_outputscale = Parameter(torch.tensor(float(outputscale), dtype=torch.float64))

    @property
    def outputscale(self) -> Tensor:
        return self.raw_outputscale

    @outputscale.setter
    def outputscale(self, value) -> None:
        with torch.no_grad():
            self.raw_outputscale.copy_(torch.as_tensor(value, dtype=torch.float64).reshape(()))


class ConstantMean(Module):
    def __init__(self, constant: float = 0.0) -> None:
        super().__init__()
        self.raw_constant = Parameter(torch.tensor(float(constant), dtype=torch.float64))

    @property
    def constant(self) -> Tensor:
        return self.raw_constant

    @constant.setter
    def constant(self, value) -> None:
        with torch.no_grad():
            self.raw_constant.copy_(torch.as_tensor(value, dtype=torch.float64).reshape(()))


class GaussianLikelihood(Module):
    """Homoskedastic noise, constrained >= MIN_INFERRED_NOISE_LEVEL (gpytorch_modules.py:74-96)."""

    def __init__(self, noise: float = math.exp(-5.0)) -> None:  # LogNormal(-4, 1) prior mode
        super().__init__()
        self.raw_noise = Parameter(torch.tensor([max(float(noise), MIN_INFERRED_NOISE_LEVEL)], dtype=torch.float64))

    @property
    def noise(self) -> Tensor:
        return self.raw_noise

    @noise.setter
    def noise(self, value) -> None:
        with torch.no_grad():
            self.raw_noise.copy_(torch.as_tensor(value, dtype=torch.float64).reshape(1).clamp_min(MIN_INFERRED_NOISE_LEVEL))


class FixedNoiseGaussianLikelihood(Module):
    """Per-observation noise `train_Yvar` (already in standardised units)."""

    def __init__(self, noise: Tensor) -> None:
        super().__init__()
        self.register_buffer("_noise", noise.reshape(-1).to(torch.float64))

    @property
    def noise(self) -> Tensor:
        return self._noise


def get_covar_module_with_dim_scaled_prior(ard_num_dims: int, use_rbf_kernel: bool = True) -> Kernel:
    """Default SingleTaskGP kernel: ARD RBF (no ScaleKernel) initialised at the mode of the
    LogNormal(sqrt2 + log(d)/2, sqrt3) lengthscale prior (gpytorch_modules.py:100-133)."""
    mode = math.exp(math.sqrt(2.0) + 0.5 * math.log(ard_num_dims) - 3.0)
    cls = RBFKernel if use_rbf_kernel else MaternKernel
    return cls(ard_num_dims=ard_num_dims, lengthscale=torch.full((ard_num_dims,), mode, dtype=torch.float64))
