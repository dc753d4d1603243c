import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("CUDA not available")
    return torch.device("cuda:0")


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests need a CUDA device AND the built extension: skip them (instead of erroring at import or at the
    first CUDA call) when either is missing, so a bare `pytest tests` is green on a CUDA-less box."""
    import torch

    from botorch_b200 import _lib

    reason = None
    if not torch.cuda.is_available():
        reason = "CUDA not available"
    elif not _lib._SO.exists():
        reason = f"{_lib._SO.name} is not built"
    if reason is None:
        return
    skip = pytest.mark.skip(reason=reason)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _restore_settings():
    """Tests flip the package-level flags; every test starts from and returns to the product defaults."""
    from botorch_b200 import settings

    saved = (settings.contraction.value(), settings.int8_slices.value(), settings.int8_cond_limit.value())
    yield
    settings.contraction.set(saved[0])
    settings.optimizer.set("scipy")
    settings.int8_slices.set(saved[1])
    settings.int8_cond_limit.set(saved[2])
    settings.int8_max_slices.set(False)
    settings.fused_log_hvi.set(True)
    settings.int8_gram.set(True)
