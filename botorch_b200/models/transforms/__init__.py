from .input import Normalize  # noqa: F401
from .outcome import Standardize  # noqa: F401
