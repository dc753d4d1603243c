from .gen import gen_candidates_scipy  # noqa: F401
