from .base import MCSampler  # noqa: F401
from .get_sampler import get_sampler  # noqa: F401
from .normal import IIDNormalSampler, NormalMCSampler, SobolQMCNormalSampler  # noqa: F401
