"""Oracle-independent anchors for the CUDA path: textbook GP formulas evaluated with plain dense linear algebra, the
interpolation and far-field limits, analytic log-EI at q = 1, and cross-checks between independent kernel modes."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _model(n=60, d=3, noise=1e-4, kernel="matern", seed=0, standardize=False):
    from botorch_b200.models import MaternKernel, RBFKernel, ScaleKernel, SingleTaskGP
    from botorch_b200.models.transforms import Standardize

    g = torch.Generator().manual_seed(seed)
    X = torch.rand(n, d, generator=g, dtype=torch.float64)
    Y = torch.cos(4 * X[:, :1]) + X[:, 1:2] ** 2 + 0.3
    ls = 0.25 + 0.3 * torch.rand(d, generator=g, dtype=torch.float64)
    base = (MaternKernel if kernel == "matern" else RBFKernel)(ard_num_dims=d, lengthscale=ls)
    model = SingleTaskGP(X.to(DEV), Y.to(DEV), covar_module=ScaleKernel(base, outputscale=1.3),
                         outcome_transform=Standardize(m=1) if standardize else None).to(DEV)
    model.likelihood.noise = noise
    return model, X, Y, ls, g


def _kern(kernel, A, B, ls, os_):
    D = ((A.unsqueeze(-2) - B.unsqueeze(-3)) / ls).pow(2).sum(-1)
    if kernel == "rbf":
        return os_ * torch.exp(-0.5 * D)
    r = D.clamp_min(0).sqrt() * math.sqrt(5)
    return os_ * (1 + r + r * r / 3) * torch.exp(-r)


@pytest.mark.parametrize("kernel", ["rbf", "matern"])
def test_posterior_equals_textbook_formula(kernel):
    model, X, Y, ls, g = _model(kernel=kernel)
    K = _kern(kernel, X, X, ls, 1.3) + 1e-4 * torch.eye(X.shape[0], dtype=torch.float64)
    Xq = torch.rand(5, 4, 3, generator=g, dtype=torch.float64)
    post = model.posterior(Xq.to(DEV))
    for b in range(5):
        Ks = _kern(kernel, Xq[b], X, ls, 1.3)
        mean = Ks @ torch.linalg.solve(K, Y.squeeze(-1))
        cov = _kern(kernel, Xq[b], Xq[b], ls, 1.3) - Ks @ torch.linalg.solve(K, Ks.T)
        assert torch.allclose(post.mean[b, :, 0].cpu(), mean, atol=1e-8)
        assert torch.allclose(post.distribution.covariance_matrix[b].cpu(), cov, atol=1e-8)


def test_interpolation_and_far_field_limits():
    model, X, Y, ls, g = _model(noise=1e-4, standardize=True)
    post = model.posterior(X[:8].unsqueeze(0).to(DEV))
    assert torch.allclose(post.mean[0, :, 0].cpu(), Y[:8, 0], atol=2e-2)        # reproduces the data ...
    s2 = float(Y.var())
    assert (post.variance[0, :, 0].cpu() < 5e-4 * s2 * 1.3 + 1e-6).all()        # ... with (almost) no uncertainty left
    far = torch.full((1, 1, 3), 50.0, dtype=torch.float64, device=DEV)           # far from all data: the prior
    pf = model.posterior(far)
    assert abs(float(pf.mean) - float(Y.mean())) < 1e-9 * max(1.0, abs(float(Y.mean())))
    assert abs(float(pf.variance) - 1.3 * s2) < 1e-9 * 1.3 * s2


def test_q1_matches_analytic_log_ei():
    """q = 1: the MC estimate with 4096 Sobol samples must agree with log(sigma * (z Phi(z) + phi(z))) computed from the
    kernel-produced posterior moments (an anchor that involves neither the oracle nor the sampling code path twice)."""
    from botorch_b200.acquisition import qLogExpectedImprovement
    from botorch_b200.sampling import SobolQMCNormalSampler

    model, X, Y, ls, g = _model(noise=1e-3)
    best = float(Y.median())
    acqf = qLogExpectedImprovement(model, best_f=best, sampler=SobolQMCNormalSampler(torch.Size([4096]), seed=0),
                                   tau_relu=1e-9)
    Xq = torch.rand(16, 1, 3, generator=g, dtype=torch.float64).to(DEV)
    with torch.no_grad():
        mc = acqf(Xq).cpu()
        post = model.posterior(Xq)
    mu, sig = post.mean.reshape(-1).cpu(), post.variance.reshape(-1).sqrt().cpu()
    z = (mu - best) / sig
    nrm = torch.distributions.Normal(0.0, 1.0)
    analytic = torch.log(sig * (z * nrm.cdf(z) + torch.exp(nrm.log_prob(z))))
    keep = z > -1.0  # where improvement is not a tail event the QMC error of 4096 points is ~1e-3
    assert keep.sum() >= 4
    assert float((mc[keep] - analytic[keep]).abs().max()) < 1e-2, (mc[keep] - analytic[keep])


def test_qlogei_tends_to_log_qei_and_modes_cross_check():
    """tau -> 0: the smoothed log-improvement path (utility mode 1, fatmax, logmeanexp) must approach log of the hard
    qEI path (utility mode 2, amax, mean) -- two independent branches of the same kernel (reference test_logei.py:128-163)."""
    from botorch_b200.acquisition import qExpectedImprovement, qLogExpectedImprovement
    from botorch_b200.sampling import SobolQMCNormalSampler

    model, X, Y, ls, g = _model(noise=1e-3)
    best = float(Y.median())
    mk = lambda: SobolQMCNormalSampler(torch.Size([512]), seed=4)
    Xq = torch.rand(12, 3, 3, generator=g, dtype=torch.float64).to(DEV)
    with torch.no_grad():
        log_qei = qExpectedImprovement(model, best_f=best, sampler=mk())(Xq).log()
        qlogei = qLogExpectedImprovement(model, best_f=best, sampler=mk(), tau_relu=1e-8, tau_max=1e-5)(Xq)
    assert float((qlogei - log_qei).abs().max()) < 1e-3
