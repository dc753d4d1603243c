from .fused_log_areas import fused_log_areas  # noqa: F401
