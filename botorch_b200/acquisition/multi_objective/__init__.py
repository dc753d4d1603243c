from .fused_log_areas import fused_log_areas  # noqa: F401
from .logei import qLogExpectedHypervolumeImprovement  # noqa: F401
