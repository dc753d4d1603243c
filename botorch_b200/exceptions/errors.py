"""Error classes at the drop-in boundary (mirrors botorch/exceptions/errors.py:16-94 and the
linear_operator errors the reference's hot path raises)."""
from __future__ import annotations


class BotorchError(Exception):
    """Base botorch exception."""


class CandidateGenerationError(BotorchError):
    """Exception raised during generating candidates."""


class DeprecationError(BotorchError):
    """Exception raised due to deprecations."""


class InputDataError(BotorchError):
    """Exception raised when input data does not comply with conventions."""


class UnsupportedError(BotorchError):
    """Currently unsupported feature."""


class BotorchTensorDimensionError(BotorchError):
    """Exception raised when a tensor violates a botorch convention."""


class ModelFittingError(Exception):
    """Exception raised when attempts to fit a model terminate unsuccessfully."""


class OptimizationTimeoutError(BotorchError):
    """Exception raised when optimization times out."""

    def __init__(self, /, *args, current_x, runtime: float, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.current_x = current_x
        self.runtime = runtime


class OptimizationGradientError(BotorchError, RuntimeError):
    """Exception raised when gradient array `gradf` contains NaNs (generation/gen.py:471-479)."""

    def __init__(self, /, *args, current_x, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.current_x = current_x


class InfeasibilityError(BotorchError, ValueError):
    """Exception raised when infeasibility occurs."""


class NanError(RuntimeError):
    """NaN/Inf encountered (linear_operator.utils.errors.NanError)."""


class NotPSDError(RuntimeError):
    """Matrix not positive definite after jitter escalation (linear_operator.utils.errors.NotPSDError)."""
