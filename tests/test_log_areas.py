"""N2: fused log-areas kernel.  CPU: the torch restatement (oracle/log_areas.py) against the REAL reference kernel
(compiled botorch/csrc/logei_fused.cpp in oracle/_ref, and golden vectors generated from it).  GPU: the CUDA op
against the same real-reference outputs.  Tolerances: the reference's own fused-vs-python test uses atol = rtol = 1e-6
(test/acquisition/multi_objective/test_logei.py:194-463); we hold 1e-11 in fp64."""
import os

import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden", "log_areas_ref.pt")


def _cases():
    return torch.load(GOLD, weights_only=False)


def _ref_module():
    from oracle.build_ref import build, load_ref

    build()
    return load_ref()


def test_restatement_matches_real_reference_golden():
    from oracle.log_areas import log_areas

    for name, c in _cases().items():
        obj = c["obj"].clone().requires_grad_(True)
        out = log_areas(obj, c["lo"], c["hi"], 1e-6, 1e-2)
        (g,) = torch.autograd.grad((out * c["go"]).sum(), obj)
        assert torch.allclose(out, c["fwd"], rtol=1e-12, atol=1e-12), name
        assert torch.allclose(g, c["bwd"], rtol=1e-9, atol=1e-9), name


def test_golden_matches_live_compiled_reference():
    ref = _ref_module()
    if ref is None:
        pytest.skip("reference source not present and oracle/_ref not prebuilt")
    for name, c in _cases().items():
        assert torch.equal(ref.forward(c["obj"], c["lo"], c["hi"], 1e-6, 1e-2), c["fwd"]), name
        assert torch.equal(ref.backward(c["go"], c["obj"], c["lo"], c["hi"], 1e-6, 1e-2), c["bwd"]), name


@pytest.mark.gpu
def test_cuda_log_areas_matches_real_reference():
    from botorch_b200.acquisition.multi_objective import fused_log_areas

    dev = torch.device("cuda:0")
    for name, c in _cases().items():
        obj = c["obj"].to(dev).requires_grad_(True)
        out = fused_log_areas(obj, c["lo"].to(dev), c["hi"].to(dev), 1e-6, 1e-2)
        (g,) = torch.autograd.grad((out * c["go"].to(dev)).sum(), obj)
        assert out.shape == c["fwd"].shape
        assert torch.allclose(out.cpu(), c["fwd"], rtol=1e-11, atol=1e-11), name
        assert torch.allclose(g.cpu(), c["bwd"], rtol=1e-9, atol=1e-10), name


@pytest.mark.gpu
def test_cuda_log_areas_grid_vs_compiled_reference_fp64_fp32():
    """Grid of (m, q-subset size, n_sub, cells) like the reference's own fused-kernel test, both dtypes."""
    from botorch_b200.acquisition.multi_objective import fused_log_areas

    ref = _ref_module()
    if ref is None:
        pytest.skip("oracle/_ref not available")
    dev = torch.device("cuda:0")
    torch.manual_seed(7)
    for dtype, tol in ((torch.float64, 1e-10), (torch.float32, 2e-3)):
        for (B, n_sub, isz, m, nc, batched) in [(64, 4, 1, 2, 16, False), (33, 6, 2, 2, 9, True), (17, 4, 3, 4, 31, True),
                                                (5, 1, 4, 8, 3, False), (3, 2, 32, 2, 2, False), (129, 10, 3, 3, 40, False)]:
            obj = torch.randn(B, n_sub, isz, m, dtype=dtype)
            lo = torch.randn(*((B,) if batched else ()), nc, m, dtype=dtype) - 0.5
            hi = lo + torch.rand_like(lo) * 2 + 0.05
            go = torch.randn(B, nc, n_sub, dtype=dtype)
            f_ref = ref.forward(obj, lo, hi, 1e-6, 1e-2)
            b_ref = ref.backward(go, obj, lo, hi, 1e-6, 1e-2)
            od = obj.to(dev).requires_grad_(True)
            f = fused_log_areas(od, lo.to(dev), hi.to(dev), 1e-6, 1e-2)
            (b,) = torch.autograd.grad((f * go.to(dev)).sum(), od)
            scale_b = b_ref.abs().max().clamp_min(1.0)
            assert torch.allclose(f.cpu(), f_ref, rtol=tol, atol=tol * 10), (dtype, B, isz, m)
            assert float((b.cpu() - b_ref).abs().max() / scale_b) < tol * 10, (dtype, B, isz, m)


@pytest.mark.gpu
def test_cuda_log_areas_full_size_properties():
    """Config-5-sized slice (S=512 MC x 64 q-batches, q=4 -> subsets of size 2, m=4, 64 cells): finite, and invariant
    under splitting B (items are independent)."""
    from botorch_b200.acquisition.multi_objective import fused_log_areas

    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    B, n_sub, isz, m, nc = 512 * 64, 6, 2, 4, 64
    obj = torch.randn(B, n_sub, isz, m, dtype=torch.float64, device=dev)
    lo = torch.randn(nc, m, dtype=torch.float64, device=dev)
    hi = lo + torch.rand_like(lo) + 0.1
    full = fused_log_areas(obj, lo, hi, 1e-6, 1e-2)
    assert torch.isfinite(full).all()
    part = torch.cat([fused_log_areas(obj[:1000], lo, hi, 1e-6, 1e-2), fused_log_areas(obj[1000:], lo, hi, 1e-6, 1e-2)])
    assert torch.equal(full, part)
