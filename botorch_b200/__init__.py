"""botorch_b200: a B200-native (sm_100a) engine for BoTorch's batched Monte-Carlo acquisition hot path.

The package mirrors the reference's module layout for the path it accelerates
(`acquisition`, `models`, `posteriors`, `sampling`, `optim`, `generation`, `utils`) and routes all
numerical work through the C-ABI CUDA library in `csrc/` (see include/mcacq_b200.h).
"""
__version__ = "0.1.0"
