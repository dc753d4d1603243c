from .gen import gen_candidates_scipy  # noqa: F401
from .sampling import MaxPosteriorSampling  # noqa: F401
from .device_gen import DeviceLBFGSB, gen_candidates_device  # noqa: F401,E402
