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
        vtol, gtol = (1e-8, 1e-6) if c["contraction"] == "dmma" else (1e-7, 2e-5)
        assert v.shape == (b,)
        assert float(((v.detach().cpu() - v_o).abs() / v_o.abs().clamp_min(1e-12)).max()) < vtol, c
        assert float((gr.cpu() - g_o).abs().max() / g_o.abs().max().clamp_min(1e-300)) < gtol, c
    finally:
        settings.contraction.set("dmma")
