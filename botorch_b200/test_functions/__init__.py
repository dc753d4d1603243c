from .synthetic import Ackley, Hartmann  # noqa: F401
