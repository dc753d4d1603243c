"""Build the REAL reference kernel `botorch/csrc/logei_fused.cpp` into oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

The reference's one native file compiles from its own single source with torch's C++ extension API (pybind11 +
ATen headers ship with torch), so it is compiled from where it lies under /root/reference -- no source is copied
into this repository.  Output: oracle/_ref/logei_fused_ref*.so (git-ignored, travels to the GPU box).  Flags: -O3
(the reference adds -march=native, acquisition/multi_objective/logei.py:93; omitted so the binary also runs on the
GPU box's host CPU).  Everything else on the hot path is Python over un-vendored gpytorch -> not buildable.
"""
from __future__ import annotations

import glob
import importlib.util
import os

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SRC = "/root/reference/botorch/csrc/logei_fused.cpp"
NAME = "logei_fused_ref"


def build(verbose: bool = False) -> str | None:
    if not os.path.exists(SRC):
        return None
    existing = glob.glob(os.path.join(OUT, NAME + "*.so"))
    if existing and os.path.getmtime(existing[0]) >= os.path.getmtime(SRC):
        return existing[0]
    os.makedirs(OUT, exist_ok=True)
    from torch.utils.cpp_extension import load

    load(name=NAME, sources=[SRC], extra_cflags=["-O3"], build_directory=OUT, verbose=verbose)
    return glob.glob(os.path.join(OUT, NAME + "*.so"))[0]


def load_ref():
    """Import the prebuilt reference extension (returns None when it was never built)."""
    hits = glob.glob(os.path.join(OUT, NAME + "*.so"))
    if not hits:
        return None
    import torch  # noqa: F401  (libtorch symbols must be loaded first)

    spec = importlib.util.spec_from_file_location(NAME, hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose=True))
