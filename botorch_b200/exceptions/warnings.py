"""Warning classes (mirrors botorch/exceptions/warnings.py:14-109)."""
from __future__ import annotations


class BotorchWarning(Warning):
    """Base botorch warning."""


class BadInitialCandidatesWarning(BotorchWarning):
    """Warning issued if set of initial candidates for optimziation is bad."""


class InputDataWarning(BotorchWarning):
    """Warning raised when input data does not comply with conventions."""


class CostAwareWarning(BotorchWarning):
    """Warning raised in the context of cost-aware acquisition strategies."""


class OptimizationWarning(BotorchWarning):
    """Optimization-related warnings."""


class SamplingWarning(BotorchWarning):
    """Sampling related warnings."""


class BotorchTensorDimensionWarning(BotorchWarning):
    """Warning raised when a tensor possibly violates a botorch convention."""


class UserInputWarning(BotorchWarning):
    """Warning raised when a potential issue is detected with user provided inputs."""


class NumericsWarning(BotorchWarning):
    """Warning raised when numerical issues are detected."""


class NumericalWarning(RuntimeWarning):
    """Jitter added to make a matrix positive definite (linear_operator NumericalWarning)."""
