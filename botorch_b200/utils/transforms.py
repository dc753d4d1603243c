"""Input-shape conventions of acquisition `forward` (reference: botorch/utils/transforms.py).

Host-side glue only: `t_batch_mode_transform` (:239-323) guarantees an explicit t-batch dimension and
checks the output shape, `concatenate_pending_points` (:378-403) appends `X_pending` along q,
`match_batch_shape` (:406-428) broadcasts batch dims, `standardize` / `normalize` / `unnormalize`
(:27-129) are the usual affine helpers.
"""
from __future__ import annotations

import functools
import warnings

import torch
from torch import Tensor


def standardize(Y: Tensor) -> Tensor:
    """Zero-mean / unit-variance per output column (reference :27-47); constant columns are only centred."""
    stddim = -1 if Y.dim() < 2 else -2
    Y_std = Y.std(dim=stddim, keepdim=True)
    Y_std = Y_std.where(Y_std >= 1e-9, torch.full_like(Y_std, 1.0))
    return (Y - Y.mean(dim=stddim, keepdim=True)) / Y_std


def normalize(X: Tensor, bounds: Tensor) -> Tensor:
    return (X - bounds[0]) / (bounds[1] - bounds[0])


def unnormalize(X: Tensor, bounds: Tensor) -> Tensor:
    return X * (bounds[1] - bounds[0]) + bounds[0]


def is_ensemble(model) -> bool:
    return bool(getattr(model, "_is_ensemble", False))


def match_batch_shape(X: Tensor, Y: Tensor) -> Tensor:
    """Expand the batch dims of X to those of Y (reference :406-428)."""
    return X.expand(X.shape[: -(Y.dim())] + Y.shape[:-2] + X.shape[-2:])


def _output_shape_ok(acqf, X: Tensor, output: Tensor) -> bool:
    xb = X.shape[:-2]
    if output.shape == xb or (output.shape == torch.Size() and xb == torch.Size([1])):
        return True
    try:
        mb = acqf.model.batch_shape
    except (AttributeError, NotImplementedError):
        warnings.warn(
            f"Output shape checks failed! Expected output shape to match t-batch shape of X, but got output with "
            f"shape {output.shape} for X with shape {X.shape}. Make sure that this is the intended behavior!",
            RuntimeWarning, stacklevel=3)
        return True
    if output.shape == mb:
        return True
    k = len(mb)
    return output.shape == xb[: len(xb) - k] + mb and all(x in (1, m) for x, m in zip(xb[len(xb) - k:], mb))


def t_batch_mode_transform(expected_q: int | None = None, assert_output_shape: bool = True):
    """Decorator factory: make `X` at least 3-dimensional (`b x q x d`) and verify the result shape."""

    def decorator(method):
        @functools.wraps(method)
        def wrapped(acqf, X, *args, **kwargs):
            if not isinstance(X, Tensor):
                return method(acqf, X, *args, **kwargs)
            if X.dim() < 2:
                raise ValueError(
                    f"{type(acqf).__name__} requires X to have at least 2 dimensions, but received X with only "
                    f"{X.dim()} dimensions.")
            if expected_q is not None and X.shape[-2] != expected_q:
                raise AssertionError(
                    f"Expected X to be `batch_shape x q={expected_q} x d`, but got X with shape {X.shape}.")
            Xb = X if X.dim() > 2 else X.unsqueeze(0)
            out = method(acqf, Xb, *args, **kwargs)
            if assert_output_shape and not _output_shape_ok(acqf, Xb, out):
                raise AssertionError(
                    "Expected the output shape to match either the t-batch shape of X, or the `model.batch_shape` "
                    f"in the case of acquisition functions using batch models; but got output with shape "
                    f"{out.shape} for X with shape {Xb.shape}.")
            return out

        return wrapped

    return decorator


def concatenate_pending_points(method):
    """Decorator: evaluate on `cat([X, X_pending], dim=-2)` when the acquisition has pending points."""

    @functools.wraps(method)
    def wrapped(acqf, X, **kwargs):
        if acqf.X_pending is not None:
            X = torch.cat([X, match_batch_shape(acqf.X_pending, X)], dim=-2)
        return method(acqf, X, **kwargs)

    return wrapped
