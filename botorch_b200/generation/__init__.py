from .gen import gen_candidates_scipy  # noqa: F401
from .sampling import MaxPosteriorSampling  # noqa: F401
