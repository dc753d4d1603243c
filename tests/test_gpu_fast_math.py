"""GPU: the kernels' own FP64 exp / log / log1p (csrc/fast_math.cuh: constant-memory coefficients instead of libdevice's
64-bit immediates) against torch's float64 functions on the device -- which is what the reference's safe_math evaluates
(utils/safe_math.py:298-355) -- over the ranges the utility kernels use, plus the special values that take the libdevice
fallback.  Bound: 3 ulp against torch (itself <= 1 ulp), i.e. ~3e-16 relative; the host-side check against long double
(tools/fast_math/check.cpp) measures 1.1 / 2.0 / 2.4 ulp."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _probe(kind, x):
    from botorch_b200 import _lib

    x = x.to(DEV, torch.float64).contiguous()
    y = torch.empty_like(x)
    rc = _lib.lib().mcacq_fast_math_probe(kind, x.data_ptr(), y.data_ptr(), x.numel(), _lib.stream_ptr())
    assert rc == 0
    torch.cuda.synchronize()
    return y


def _ulps(got, want):
    spacing = torch.ldexp(torch.ones_like(want), torch.frexp(want)[1] - 53)
    return ((got - want).abs() / spacing).max().item()


def test_exp_log_log1p_within_3_ulp_of_torch():
    g = torch.Generator().manual_seed(0)
    n = 1 << 21
    u = torch.rand(n, generator=g, dtype=torch.float64)
    x = torch.cat([(u - 0.5) * 1416.0, (u - 0.5) * 4.0 * 10.0 ** (-12.0 * torch.rand(n, generator=g, dtype=torch.float64))])
    assert _ulps(_probe(0, x), torch.exp(x.to(DEV))) <= 3.0
    mant = 1.0 + torch.rand(n, generator=g, dtype=torch.float64)
    ex = torch.randint(-1022, 1024, (n,), generator=g)
    x = torch.cat([torch.ldexp(mant, ex), 1.0 + (u - 0.5) * 10.0 ** (-15.0 * torch.rand(n, generator=g, dtype=torch.float64)),
                   1.0 + 31.0 * u])
    want = torch.log(x.to(DEV))
    keep = want != 0
    assert _ulps(_probe(1, x)[keep], want[keep]) <= 3.0
    x = torch.exp(-660.0 + 680.0 * u)
    assert _ulps(_probe(2, x), torch.log1p(x.to(DEV))) <= 3.5
    tiny = torch.exp(-744.0 + 80.0 * u[:4096])           # down into the denormals: log1p(x) == x exactly
    assert torch.equal(_probe(2, tiny).cpu(), tiny)


def test_special_values_follow_libdevice():
    inf, nan = math.inf, math.nan
    x = torch.tensor([-800.0, -745.0, -708.5, 708.5, 709.7, 800.0, inf, -inf, nan, 0.0, -0.0], dtype=torch.float64)
    got, want = _probe(0, x).cpu(), torch.exp(x.to(DEV)).cpu()   # (libdevice flushes exp(-745) to 0: compare on the device)
    assert torch.equal(got.isnan(), want.isnan())
    assert torch.allclose(got[~want.isnan()], want[~want.isnan()], rtol=4e-16, atol=0)
    x = torch.tensor([0.0, -0.0, -1.0, 5e-324, 2e-308, 1.0, inf, nan, 1.7e308], dtype=torch.float64)
    got, want = _probe(1, x).cpu(), torch.log(x.to(DEV)).cpu()
    assert torch.equal(got.isnan(), want.isnan())
    assert torch.allclose(got[~want.isnan()], want[~want.isnan()], rtol=4e-16, atol=0)
    assert float(got[5]) == 0.0
    x = torch.tensor([0.0, 1e-320, 1e-300, 1.0, 1e300, 1e305, inf, nan], dtype=torch.float64)
    got, want = _probe(2, x).cpu(), torch.log1p(x.to(DEV)).cpu()
    assert torch.equal(got.isnan(), want.isnan())
    assert torch.allclose(got[~want.isnan()], want[~want.isnan()], rtol=5e-16, atol=0)
