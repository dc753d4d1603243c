from .gpytorch import GPyTorchPosterior, MultivariateNormal  # noqa: F401
from .posterior import Posterior  # noqa: F401
