from .errors import (  # noqa: F401
    BotorchError, BotorchTensorDimensionError, CandidateGenerationError, InputDataError, InfeasibilityError,
    ModelFittingError, NanError, NotPSDError, OptimizationGradientError, OptimizationTimeoutError, UnsupportedError,
)
from .warnings import (  # noqa: F401
    BadInitialCandidatesWarning, BotorchTensorDimensionWarning, BotorchWarning, InputDataWarning, NumericalWarning,
    NumericsWarning, OptimizationWarning, SamplingWarning, UserInputWarning,
)
