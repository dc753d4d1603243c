"""MC sampler base (reference: botorch/sampling/base.py:32-150): holds `sample_shape`, `seed` and the
cached `base_samples` buffer whose t-batch dimensions are collapsed to size 1."""
from __future__ import annotations

from abc import ABC, abstractmethod

import torch
from torch import Tensor
from torch.nn import Module

from ..exceptions.errors import InputDataError


class MCSampler(Module, ABC):
    def __init__(self, sample_shape: torch.Size, seed: int | None = None) -> None:
        super().__init__()
        if not isinstance(sample_shape, torch.Size):
            raise InputDataError(f"Expected `sample_shape` to be a `torch.Size` object, got {sample_shape}.")
        self.sample_shape = sample_shape
        # un-seeded samplers draw their seed from torch's global RNG, as the reference does (:64-66)
        self.seed = seed if seed is not None else int(torch.randint(0, 1000000, (1,)).item())
        self.register_buffer("base_samples", None)

    @abstractmethod
    def forward(self, posterior) -> Tensor:
        ...

    def _get_batch_range(self, posterior) -> tuple[int, int]:
        if hasattr(self, "batch_range_override"):
            return self.batch_range_override
        return posterior.batch_range

    def _get_collapsed_shape(self, posterior) -> torch.Size:
        bss = posterior.base_sample_shape
        lo, hi = self._get_batch_range(posterior)
        collapsed = bss[:lo] + torch.Size([1 for _ in bss[lo:hi]]) + bss[hi:]
        return self.sample_shape + collapsed

    def _get_extended_base_sample_shape(self, posterior) -> torch.Size:
        return self.sample_shape + posterior.base_sample_shape

    def _update_base_samples(self, posterior, base_sampler: "MCSampler") -> None:
        raise NotImplementedError(f"{self.__class__.__name__} does not implement `_update_base_samples`.")

    def _instance_check(self, base_sampler) -> None:
        if not isinstance(base_sampler, self.__class__):
            raise RuntimeError(
                f"Expected `base_sampler` to be an instance of {self.__class__.__name__}. Got {base_sampler}.")
