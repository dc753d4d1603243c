"""GPU: qLogNEI with baselines of MORE than 64 points on the fused route (missing #8: r > 64 used to take the
generic torch-op route).  The posterior-block kernels sweep such a baseline in 64-row chunks (forward: one launch per chunk,
backward: a run-time loop over the baseline rows); the sample / reduce kernels are generic in r.  Reference semantics:
`sample_cached_cholesky` (botorch/utils/low_rank.py:84-172) with `baseline_L` of the whole baseline
(botorch/acquisition/cached_cholesky.py:98-124).  Tolerances: value 1e-9, gradient 1e-7 (BASELINE.json north_star)."""
import os
from dataclasses import replace

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = torch.device("cuda:0")


def _setup(cfg, n, r, q, S=256, b=12):
    from botorch_b200.benchmarks import configs
    from oracle.harness import build_oracle

    spec = replace(configs.CONFIGS[cfg], r=r, q=q, S=S)
    data = configs.make_problem(spec, n=n)
    model = configs.build_model(data, DEV)
    acqf = configs.build_acqf(data, model)
    return data, model, acqf, build_oracle(data), configs.eval_points(data, b)


def test_fused_supported_query():
    from botorch_b200 import _lib

    assert _lib.fused_supported(8, 16, 1024) and _lib.fused_supported(8, 64, 1024)
    assert _lib.fused_supported(8, 65, 1024) and _lib.fused_supported(8, 512, 1024)
    assert not _lib.fused_supported(8, 513, 1024)       # beyond MCACQ_MAX_R
    assert not _lib.fused_supported(33, 16, 1024)       # beyond MCACQ_MAX_Q
    assert not _lib.fused_supported(32, 512, 1024)      # compiled limits fine, shared memory of sample / reduce is not


@pytest.mark.parametrize("mode", ["int8", "dmma"])
@pytest.mark.parametrize("cfg,n,r,q", [("C2", 300, 65, 8), ("C3", 400, 100, 3), ("C3", 512, 200, 8), ("C2", 384, 130, 12),
                                       ("C3", 640, 300, 20)])
def test_value_and_gradient_match_the_oracle(cfg, n, r, q, mode):
    from botorch_b200 import settings
    from botorch_b200.acquisition import logei as _fused
    from oracle.acquisition import value_and_grad

    with settings.contraction(mode):
        data, model, acqf, orc, X = _setup(cfg, n, r, q)
        assert acqf.X_baseline.shape[-2] == r
        calls = []
        orig = _fused.fused_acquisition
        _fused.fused_acquisition = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
        try:
            Xg = X.to(DEV).requires_grad_(True)
            v = acqf(Xg)
            (g,) = torch.autograd.grad(v.sum(), Xg)
        finally:
            _fused.fused_acquisition = orig
    assert calls, "the large baseline did not take the fused route"
    v_o, g_o = value_and_grad(orc, X)
    assert float(((v.detach().cpu() - v_o).abs() / v_o.abs()).max()) < 1e-9
    assert float((g.cpu() - g_o).abs().max() / g_o.abs().max()) < 1e-7


@pytest.mark.parametrize("mode", ["int8", "dmma"])
@pytest.mark.parametrize("r,q", [(16, 8), (40, 8), (64, 12), (24, 20)])
def test_runtime_baseline_loop_is_bit_identical_to_the_unrolled_backward(r, q, mode):
    """Same DMMA sequence per output element: forcing the r > 64 backward kernel onto r <= 64 must not change one bit."""
    from botorch_b200 import settings

    with settings.contraction(mode):
        data, model, acqf, orc, X = _setup("C3", 320, r, q)
        grads = []
        for force in ("0", "1"):
            os.environ["MCACQ_BWD_RLOOP"] = force
            try:
                Xg = X.to(DEV).requires_grad_(True)
                (g,) = torch.autograd.grad(acqf(Xg).sum(), Xg)
                grads.append(g)
            finally:
                os.environ.pop("MCACQ_BWD_RLOOP", None)
    assert torch.equal(grads[0], grads[1])


def test_chunked_forward_equals_the_generic_route():
    """Fused (chunked baseline) vs the generic route over `model.posterior(cat[X_baseline, X])` (the reference's
    `test_cache_root` analogue, test/acquisition/test_logei.py:513-629)."""
    from botorch_b200.acquisition import GenericMCObjective, qLogNoisyExpectedImprovement
    from botorch_b200.sampling import SobolQMCNormalSampler

    data, model, acqf, orc, X = _setup("C3", 400, 90, 4, b=6)
    generic = qLogNoisyExpectedImprovement(
        model, X_baseline=data.X_baseline.to(DEV), prune_baseline=False,
        sampler=SobolQMCNormalSampler(torch.Size([data.spec.S]), seed=1234),
        objective=GenericMCObjective(lambda samples, X=None: samples.squeeze(-1)))
    Xa = X.to(DEV).requires_grad_(True)
    Xb = X.to(DEV).requires_grad_(True)
    va, vb = acqf(Xa), generic(Xb)
    ga, = torch.autograd.grad(va.sum(), Xa)
    gb, = torch.autograd.grad(vb.sum(), Xb)
    assert float(((va - vb).abs() / vb.abs()).max().detach()) < 1e-9
    assert float((ga - gb).abs().max() / gb.abs().max()) < 1e-7


@pytest.mark.parametrize("mode", ["int8", "dmma"])
def test_baseline_gemm_and_chunked_sweep_agree(mode):
    """r > 64 runs the cross term / the baseline term of dA as plain GEMMs (dgemm_nt); MCACQ_BIGR_GEMM=0 selects the chunked
    sweep / the in-kernel loop instead.  Different summation orders of the same fp64 sums: agreement to rounding."""
    from botorch_b200 import settings

    with settings.contraction(mode):
        data, model, acqf, orc, X = _setup("C3", 512, 160, 8)
        out = []
        for flag in ("1", "0"):
            os.environ["MCACQ_BIGR_GEMM"] = flag
            try:
                Xg = X.to(DEV).requires_grad_(True)
                v = acqf(Xg)
                (g,) = torch.autograd.grad(v.sum(), Xg)
                out.append((v.detach(), g))
            finally:
                os.environ.pop("MCACQ_BIGR_GEMM", None)
    assert float(((out[0][0] - out[1][0]).abs() / out[1][0].abs()).max()) < 1e-11
    assert float((out[0][1] - out[1][1]).abs().max() / out[1][1].abs().max()) < (1e-9 if mode == "dmma" else 1e-7)
