from .gp_regression import SingleTaskGP  # noqa: F401
from .kernels import (ConstantMean, FixedNoiseGaussianLikelihood, GaussianLikelihood, MaternKernel, RBFKernel,  # noqa: F401
                      ScaleKernel)
from .model import Model  # noqa: F401
from .model_list_gp_regression import ModelListGP  # noqa: F401
