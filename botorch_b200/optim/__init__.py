from .initializers import (gen_batch_initial_conditions, initialize_q_batch, initialize_q_batch_nonneg,  # noqa: F401
                           initialize_q_batch_topn, is_nonnegative, sample_points_around_best)
from .optimize import optimize_acqf  # noqa: F401
