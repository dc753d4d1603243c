from .acquisition import AcquisitionFunction, MCSamplerMixin  # noqa: F401
from .logei import qLogExpectedImprovement, qLogNoisyExpectedImprovement  # noqa: F401
from .monte_carlo import MCAcquisitionFunction, SampleReducingMCAcquisitionFunction  # noqa: F401
from .objective import (GenericMCObjective, IdentityMCObjective, LinearMCObjective, MCAcquisitionObjective,  # noqa: F401
                        PosteriorTransform)
from .mc_improvement import (qExpectedImprovement, qLowerConfidenceBound, qNoisyExpectedImprovement,  # noqa: F401
                             qPosteriorStandardDeviation, qProbabilityOfImprovement, qSimpleRegret,
                             qUpperConfidenceBound)
