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


def get_X_baseline(acq_function) -> Tensor | None:
    """The acquisition function's `X_baseline`, else the model's (untransformed) training inputs
    (reference optim/utils/acquisition_utils.py:129-170)."""
    import warnings

    from ..exceptions.warnings import BotorchWarning

    try:
        X = acq_function.X_baseline
        if X.shape[0] == 0:
            raise BotorchError
    except (BotorchError, AttributeError):
        try:
            model = acq_function.model
        except AttributeError:
            warnings.warn("Failed to extract X_baseline.", BotorchWarning)
            return None
        try:
            m = model.models[0] if hasattr(model, "models") else model
            X = m.train_X_raw if hasattr(m, "train_X_raw") else m.train_inputs[0]
        except (BotorchError, AttributeError, IndexError):
            warnings.warn("Failed to extract X_baseline.", BotorchWarning)
            return None
    while X.ndim > 2:
        X = X[0]
    return X
