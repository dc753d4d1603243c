from .initializers import gen_batch_initial_conditions, initialize_q_batch, initialize_q_batch_topn  # noqa: F401
from .optimize import optimize_acqf  # noqa: F401
