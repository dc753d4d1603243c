"""CPU: the FP64 exp / log / log1p of the utility kernels (botorch_b200/csrc/fast_math.cuh) compiled for the HOST -- the same
source; only the 20-bit hardware reciprocal seed is modelled -- against long double over dense random samples of the ranges the
kernels use (tools/fast_math/check.cpp).  Bounds: 1.5 / 2.5 / 3.0 ulp (measured 1.1 / 2.0 / 2.4); the device-side check
against torch is tests/test_gpu_fast_math.py."""
import re
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_fast_math_against_long_double(tmp_path):
    exe = tmp_path / "fm_check"
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-o", str(exe), str(ROOT / "tools" / "fast_math" / "check.cpp")],
                   check=True)
    out = subprocess.run([str(exe), "2000000"], check=True, capture_output=True, text=True).stdout
    exp_ulp = float(re.search(r"fm_exp\s+max ([0-9.]+) ulp", out).group(1))
    log_ulps = [float(v) for v in re.search(r"fm_log\s+max ([0-9.]+) ulp \(normal range\), ([0-9.]+) \(near 1\), ([0-9.]+)", out).groups()]
    l1p_ulp = float(re.search(r"fm_log1p_nonneg max ([0-9.]+) ulp", out).group(1))
    assert exp_ulp <= 1.5 and max(log_ulps) <= 2.5 and l1p_ulp <= 3.0, out
    assert "exp(-800)=0 exp(800)=inf exp(nan)=nan log(0)=-inf" in out and "log(inf)=inf" in out and "log1p(0)=0" in out
