"""Small helpers used by the candidate generators (reference: botorch/optim/utils/acquisition_utils.py:22-126,
botorch/optim/parameter_constraints.py:237-246)."""
from __future__ import annotations

import numpy as np
import torch
from torch import Tensor

from ..exceptions.errors import BotorchError


def columnwise_clamp(X: Tensor, lower=None, upper=None, raise_on_violation: bool = False) -> Tensor:
    """Clamp the last dimension of X to [lower, upper] (floats or d-dim tensors)."""
    if lower is None and upper is None:
        return X
    lo = None if lower is None else torch.as_tensor(lower).expand(X.shape[-1]).to(X)
    hi = None if upper is None else torch.as_tensor(upper).expand(X.shape[-1]).to(X)
    if lo is not None and hi is not None and (lo > hi).any():
        raise ValueError("Lower bounds cannot exceed upper bounds.")
    out = X
    if lo is not None:
        out = torch.max(out, lo)
    if hi is not None:
        out = torch.min(out, hi)
    if raise_on_violation and not X.allclose(out):
        raise BotorchError("Original value(s) are out of bounds.")
    return out


def _arrayify(X: Tensor) -> np.ndarray:
    return X.cpu().detach().contiguous().double().clone().numpy()
