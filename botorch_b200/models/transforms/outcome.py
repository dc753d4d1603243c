"""`Standardize` outcome transform (reference: botorch/models/transforms/outcome.py:236-511).
Affine, so the CUDA posterior kernel folds the un-standardisation in: mean <- m + s*mean, cov <- s^2*cov."""
from __future__ import annotations

import torch
from torch import Tensor
from torch.nn import Module


class Standardize(Module):
    def __init__(self, m: int = 1, min_stdv: float = 1e-8) -> None:
        super().__init__()
        if m != 1:
            raise NotImplementedError("botorch_b200 supports single-output models (m=1) on the hot path.")
        self._m = m
        self._min_stdv = min_stdv
        self.register_buffer("means", torch.zeros(1, m, dtype=torch.float64))
        self.register_buffer("stdvs", torch.ones(1, m, dtype=torch.float64))
        self.register_buffer("_stdvs_sq", torch.ones(1, m, dtype=torch.float64))
        self._is_trained = False

    def forward(self, Y: Tensor, Yvar: Tensor | None = None, X: Tensor | None = None):
        if self.training:
            if Y.shape[-2] < 1:
                raise ValueError(f"Can't standardize with no observations. {Y.shape=}.")
            if Y.shape[-2] == 1:
                stdvs = torch.ones((1, Y.shape[-1]), dtype=Y.dtype, device=Y.device)
            else:
                # nanstd (models/transforms/utils.py:146-160): sqrt(mean((Y - mean)^2) * n / (n - 1))
                nobs = Y.shape[-2]
                stdvs = ((Y - Y.mean(dim=-2, keepdim=True)).pow(2).mean(dim=-2, keepdim=True) * nobs / (nobs - 1)).sqrt()
            stdvs = stdvs.where(stdvs >= self._min_stdv, torch.full_like(stdvs, 1.0))
            self.means = Y.mean(dim=-2, keepdim=True)
            self.stdvs = stdvs
            self._stdvs_sq = stdvs.pow(2)
            self._is_trained = True
        Y_tf = (Y - self.means) / self.stdvs
        Yvar_tf = Yvar / self._stdvs_sq if Yvar is not None else None
        return Y_tf, Yvar_tf

    def untransform(self, Y: Tensor, Yvar: Tensor | None = None):
        Y_utf = self.means + self.stdvs * Y
        return Y_utf, (Yvar * self._stdvs_sq if Yvar is not None else None)
