"""Seeded random sweep of the drop-in surface against the oracle: shapes (n, d, q, r, S, b), kernel family, ScaleKernel,
Normalize on a shifted box, fixed vs inferred noise, fat / non-fat, both contraction modes.  Cases whose q x q conditional
covariance is close to singular (the oracle's own Cholesky pivots below 1e-7 of the prior variance) are skipped: there
the comparison would measure rounding-level jitter decisions, not the kernels."""
import os
import warnings

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
N_CASES = int(os.environ.get("MCACQ_FUZZ_CASES", "80"))


def _case(seed: int):
    g = torch.Generator().manual_seed(1000 + seed)
    ri = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=g))
    n, d = ri(2, 260), ri(1, 24)
    q = ri(1, 6) if seed % 5 else ri(7, 32)
    r = 0 if seed % 3 == 0 else (ri(1, 24) if seed % 7 else ri(25, 64))
    S, b = ri(1, 300), ri(1, 23)
    if seed % 11 == 5 and r > 0:   # baselines beyond 64 points: the GEMM (even r) / run-time loop (odd r) posterior-block routes
        r, q = ri(65, 220), min(q, 12)
    return dict(n=n, d=d, q=q, r=r, S=S, b=b, kernel="rbf" if seed % 2 else "matern52", scale=seed % 4 == 1,
                normalize=seed % 3 == 1, fixed_noise=seed % 5 == 2, fat=seed % 6 != 3,
                contraction="int8" if seed % 2 == 0 else "dmma", g=g)


@pytest.mark.parametrize("seed", range(N_CASES))
def test_random_configuration_matches_oracle(seed):
    from botorch_b200 import settings
    from botorch_b200.acquisition import qLogExpectedImprovement, qLogNoisyExpectedImprovement
    from botorch_b200.models import MaternKernel, RBFKernel, ScaleKernel, SingleTaskGP
    from botorch_b200.models.transforms import Normalize
    from botorch_b200.sampling import SobolQMCNormalSampler
    from oracle.acquisition import OracleQLogEI, OracleQLogNEI, value_and_grad
    from oracle.gp import OracleGP, psd_safe_cholesky

    c = _case(seed)
    g, n, d, q, r, S, b = c["g"], c["n"], c["d"], c["q"], c["r"], c["S"], c["b"]
    lo, hi = (-2.0, 3.0) if c["normalize"] else (0.0, 1.0)
    X = lo + (hi - lo) * torch.rand(n, d, generator=g, dtype=torch.float64)
    Y = torch.sin(2.5 * (X - lo).sum(-1, keepdim=True) / (hi - lo) / d ** 0.5) + 0.1 * torch.randn(n, 1, generator=g, dtype=torch.float64)
    ls = (0.15 + 0.25 * torch.rand(d, generator=g, dtype=torch.float64)) * d ** 0.5
    os_ = 1.7 if c["scale"] else None
    bounds = torch.tensor([[lo] * d, [hi] * d], dtype=torch.float64)
    noise = (2e-3 + 5e-3 * torch.rand(n, generator=g, dtype=torch.float64)) if c["fixed_noise"] else torch.tensor(4e-3, dtype=torch.float64)
    base = (RBFKernel if c["kernel"] == "rbf" else MaternKernel)(ard_num_dims=d, lengthscale=ls)
    try:
        settings.contraction.set(c["contraction"])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            s_y = float(Y.std()) if n > 1 else 1.0
            model = SingleTaskGP(X.to(DEV), Y.to(DEV),
                                 train_Yvar=(noise * s_y**2).unsqueeze(-1).to(DEV) if c["fixed_noise"] else None,
                                 covar_module=ScaleKernel(base, os_) if os_ else base,
                                 input_transform=Normalize(d=d, bounds=bounds.to(DEV)) if c["normalize"] else None).to(DEV)
            if not c["fixed_noise"]:
                model.likelihood.noise = float(noise)
        gp = OracleGP(X, Y, ls, noise, kernel=c["kernel"], outputscale=os_,
                      norm_offset=bounds[0] if c["normalize"] else None,
                      norm_coef=(bounds[1] - bounds[0]) if c["normalize"] else None)
        Xq = lo + (hi - lo) * torch.rand(b, q, d, generator=g, dtype=torch.float64)
        sampler = SobolQMCNormalSampler(torch.Size([S]), seed=seed)
        if r == 0:
            best = Y.max() - 0.2
            acqf = qLogExpectedImprovement(model, best_f=best.to(DEV), sampler=sampler, fat=c["fat"])
            orc = OracleQLogEI(gp, best, S, seed, fat=c["fat"])
            Xall = Xq
        else:
            Xb = lo + (hi - lo) * torch.rand(r, d, generator=g, dtype=torch.float64)
            acqf = qLogNoisyExpectedImprovement(model, X_baseline=Xb.to(DEV), prune_baseline=False, sampler=sampler, fat=c["fat"])
            orc = OracleQLogNEI(gp, Xb, S, seed, fat=c["fat"])
            Xall = torch.cat([Xb.expand(b, r, d), Xq], dim=-2)
        # conditioning screen on the oracle side
        _, cov = gp.posterior_mvn(Xall)
        with warnings.catch_warnings(record=True) as ws:
            warnings.simplefilter("always")
            L = psd_safe_cholesky(cov, max_tries=6)
        prior = float(cov.diagonal(dim1=-1, dim2=-2).max())
        if ws or float(L.diagonal(dim1=-1, dim2=-2).min()) ** 2 < 1e-7 * prior:
            pytest.skip("nearly singular joint covariance: jitter decisions are rounding-level")
        v_o, g_o = value_and_grad(orc, Xq)
        Xg = Xq.to(DEV).requires_grad_(True)
        v = acqf(Xg)
        (gr,) = torch.autograd.grad(v.sum(), Xg)
        vtol, gtol = 1e-8, 1e-6   # the same bar in both contraction modes
        assert v.shape == (b,)
        assert float(((v.detach().cpu() - v_o).abs() / v_o.abs().clamp_min(1e-12)).max()) < vtol, c
        assert float((gr.cpu() - g_o).abs().max() / g_o.abs().max().clamp_min(1e-300)) < gtol, c
    finally:
        settings.contraction.set("dmma")


@pytest.mark.parametrize("seed", range(24))
def test_random_posterior_and_batch_shapes(seed):
    """`model.posterior` on multi-dimensional t-batches, non-contiguous / expanded inputs and 2-D `q x d` inputs: mean and
    full covariance against the oracle, gradients against the oracle's autograd; acquisition values keep the batch shape."""
    from botorch_b200.acquisition import qLogExpectedImprovement
    from botorch_b200.models import MaternKernel, RBFKernel, SingleTaskGP
    from botorch_b200.sampling import SobolQMCNormalSampler
    from oracle.acquisition import OracleQLogEI
    from oracle.gp import OracleGP

    g = torch.Generator().manual_seed(5000 + seed)
    ri = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=g))
    n, d, q = ri(3, 150), ri(1, 12), ri(1, 9)
    kernel = "rbf" if seed % 2 else "matern52"
    X = torch.rand(n, d, generator=g, dtype=torch.float64)
    Y = torch.cos(3 * X.sum(-1, keepdim=True) / d ** 0.5) + 0.05 * torch.randn(n, 1, generator=g, dtype=torch.float64)
    ls = (0.2 + 0.2 * torch.rand(d, generator=g, dtype=torch.float64)) * d ** 0.5
    model = SingleTaskGP(X.to(DEV), Y.to(DEV), covar_module=(RBFKernel if kernel == "rbf" else MaternKernel)(ard_num_dims=d, lengthscale=ls)).to(DEV)
    model.likelihood.noise = 3e-3
    gp = OracleGP(X, Y, ls, torch.tensor(3e-3, dtype=torch.float64), kernel=kernel)
    shape = [(), (ri(1, 4),), (ri(1, 3), ri(1, 3)), (2, 1, ri(1, 3))][seed % 4]
    Xq = torch.rand(*shape, q, d, generator=g, dtype=torch.float64)
    if seed % 3 == 0 and len(shape) >= 1:  # non-contiguous view of a larger tensor
        big = torch.rand(*shape, q, 2 * d, generator=g, dtype=torch.float64)
        Xq = big[..., ::2]
    Xo = Xq.clone().requires_grad_(True)
    m_o, c_o = gp.posterior_mvn(Xo if Xo.dim() > 2 else Xo.unsqueeze(0))
    w = torch.randn(c_o.shape, generator=g, dtype=torch.float64)
    (g_o,) = torch.autograd.grad((c_o * w).sum() + m_o.sum(), Xo)
    Xg = Xq.to(DEV).requires_grad_(True)
    post = model.posterior(Xg)
    assert post.mean.shape == (*shape, q, 1)
    cov = post.distribution.covariance_matrix
    assert float((post.mean.squeeze(-1).detach().cpu().reshape(m_o.shape) - m_o.detach()).abs().max() / m_o.detach().abs().max()) < 1e-9
    assert float((cov.detach().cpu().reshape(c_o.shape) - c_o.detach()).abs().max() / c_o.detach().abs().max()) < 1e-9
    (gr,) = torch.autograd.grad((cov.reshape(c_o.shape) * w.to(DEV)).sum() + post.mean.sum(), Xg)
    assert float((gr.cpu() - g_o).abs().max() / g_o.abs().max()) < 1e-7
    acqf = qLogExpectedImprovement(model, best_f=float(Y.max()) - 0.1, sampler=SobolQMCNormalSampler(torch.Size([32]), seed=seed))
    with torch.no_grad():
        v = acqf(Xq.to(DEV))
    # (a python-float best_f becomes a float32 buffer in the reference, `torch.as_tensor(best_f)`: same on both sides)
    v_o = OracleQLogEI(gp, float(Y.max()) - 0.1, 32, seed)(Xq.reshape(-1, q, d)).detach()
    assert v.shape == (torch.Size(shape) if shape else torch.Size([1]))  # a 2-D `q x d` input is a t-batch of one
    assert float(((v.cpu().reshape(-1) - v_o).abs() / v_o.abs().clamp_min(1e-12)).max()) < 1e-8


@pytest.mark.parametrize("seed", range(20))
def test_random_utility_modes_and_pending_points(seed):
    """qEI / qSimpleRegret / qPI / qNEI (utility modes 2-4 of the fused kernels) on random shapes, with pending points
    appended along q, against the oracle; values relative to the largest value, gradients to the largest gradient."""
    from botorch_b200.acquisition import (qExpectedImprovement, qNoisyExpectedImprovement, qProbabilityOfImprovement,
                                          qSimpleRegret)
    from botorch_b200.models import MaternKernel, RBFKernel, SingleTaskGP
    from botorch_b200.sampling import SobolQMCNormalSampler
    from oracle import acquisition as oa
    from oracle.gp import OracleGP

    g = torch.Generator().manual_seed(9000 + seed)
    ri = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=g))
    n, d, q, S, b = ri(5, 120), ri(1, 10), ri(1, 7), ri(8, 200), ri(1, 12)
    kernel = "rbf" if seed % 2 else "matern52"
    X = torch.rand(n, d, generator=g, dtype=torch.float64)
    Y = torch.sin(3 * X.sum(-1, keepdim=True) / d ** 0.5) + 0.1 * torch.randn(n, 1, generator=g, dtype=torch.float64)
    ls = (0.2 + 0.2 * torch.rand(d, generator=g, dtype=torch.float64)) * d ** 0.5
    model = SingleTaskGP(X.to(DEV), Y.to(DEV), covar_module=(RBFKernel if kernel == "rbf" else MaternKernel)(ard_num_dims=d, lengthscale=ls)).to(DEV)
    model.likelihood.noise = 5e-3
    gp = OracleGP(X, Y, ls, torch.tensor(5e-3, dtype=torch.float64), kernel=kernel)
    best = Y.median()
    npend = seed % 3  # 0, 1 or 2 pending points
    P = torch.rand(npend, d, generator=g, dtype=torch.float64) if npend else None
    Xq = torch.rand(b, q, d, generator=g, dtype=torch.float64)
    Xfull = Xq if P is None else torch.cat([Xq, P.expand(b, npend, d)], dim=-2)
    mk = lambda: SobolQMCNormalSampler(torch.Size([S]), seed=seed)
    ei = oa.OracleQLogEI(gp, best, S, seed)
    kind = seed % 4
    if kind == 0:
        acqf, ofn = qExpectedImprovement(model, best_f=best.to(DEV), sampler=mk()), lambda x: oa.oracle_qei(ei, x)
    elif kind == 1:
        acqf, ofn = qSimpleRegret(model, sampler=mk()), lambda x: oa.oracle_qsr(ei, x)
    elif kind == 2:
        acqf, ofn = (qProbabilityOfImprovement(model, best_f=best.to(DEV), sampler=mk(), tau=5e-2),
                     lambda x: oa.oracle_qpi(ei, x, tau=5e-2))
    else:
        Xb = torch.rand(ri(1, 12), d, generator=g, dtype=torch.float64)
        nei = oa.OracleQLogNEI(gp, Xb, S, seed)
        acqf, ofn = (qNoisyExpectedImprovement(model, X_baseline=Xb.to(DEV), prune_baseline=False, sampler=mk()),
                     lambda x: oa.oracle_qnei(nei, x))
    if P is not None:
        acqf.set_X_pending(P.to(DEV))
    Xo = Xq.clone().requires_grad_(True)
    v_o = ofn(Xo if P is None else torch.cat([Xo, P.expand(b, npend, d)], dim=-2))
    g_o = torch.autograd.grad(v_o.sum(), Xo)[0] if v_o.requires_grad and float(v_o.abs().sum()) > 0 else torch.zeros_like(Xq)
    Xg = Xq.to(DEV).requires_grad_(True)
    v = acqf(Xg)
    gr = torch.autograd.grad(v.sum(), Xg)[0]
    vs = float(v_o.detach().abs().max().clamp_min(1e-12))
    assert float((v.detach().cpu() - v_o.detach()).abs().max()) / vs < 1e-8
    gs = float(g_o.abs().max())
    if gs > 0:
        assert float((gr.cpu() - g_o).abs().max()) / gs < 1e-6
    assert Xfull.shape[-2] == q + npend
