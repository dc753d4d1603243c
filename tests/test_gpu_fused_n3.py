"""GPU: the round-2 fused modes of `sample_reduce` -- affine objective (`LinearMCObjective` on
the single outcome), smoothed outcome constraints evaluated inside the kernel, and the MC-mean utilities qUCB / qLCB /
qPSTD -- against the generic torch route (which tests/test_gpu_mc_utilities.py and tests/test_gpu_constraints.py pin to the
oracle) and directly against a plain-torch restatement of the reference formulas on the oracle's posterior.  Values 1e-9,
gradients 1e-7.  Every case also asserts that the fused kernels really ran (launch counter)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _setup(cfg="C2", n=192, b=12, **over):
    from dataclasses import replace

    from botorch_b200.benchmarks import configs

    spec = replace(configs.CONFIGS[cfg], **over)
    data = configs.make_problem(spec, n=n)
    model = configs.build_model(data, DEV)
    return spec, data, model, configs.eval_points(data, b).to(DEV)


def _sampler(S):
    from botorch_b200.sampling import SobolQMCNormalSampler

    return SobolQMCNormalSampler(torch.Size([S]), seed=1234)


def _vg(acqf, X):
    Xg = X.clone().requires_grad_(True)
    v = acqf(Xg)
    (g,) = torch.autograd.grad(v.sum(), Xg)
    return v.detach(), g


def _same(fused, generic, X, vtol=1e-9, gtol=1e-7):
    from botorch_b200.acquisition._fused import LaunchStats

    n0 = LaunchStats.launches
    vf, gf = _vg(fused, X)
    assert LaunchStats.launches > n0, "the fused kernels did not run"
    n1 = LaunchStats.launches
    vg, gg = _vg(generic, X)
    assert LaunchStats.launches == n1, "the reference arm of this test must be the generic torch route"
    scale = vg.abs().max().clamp_min(1e-300)
    assert float((vf - vg).abs().max() / scale) < vtol, (vf, vg)
    assert float((gf - gg).abs().max() / gg.abs().max().clamp_min(1e-300)) < gtol


def _generic_objective(w):
    from botorch_b200.acquisition.objective import GenericMCObjective

    return GenericMCObjective(lambda Y, X=None: w * Y[..., 0])


@pytest.mark.parametrize("w", [2.5, -0.7])
def test_linear_objective_on_the_fused_route(w):
    from botorch_b200.acquisition import qLogExpectedImprovement, qLogNoisyExpectedImprovement
    from botorch_b200.acquisition.objective import LinearMCObjective

    spec, data, model, X = _setup(S=128)
    lin = LinearMCObjective(torch.tensor([w], dtype=torch.float64, device=DEV))
    best = torch.tensor(w * float(data.train_Y.median()), dtype=torch.float64, device=DEV)
    _same(qLogExpectedImprovement(model, best_f=best, sampler=_sampler(128), objective=lin),
          qLogExpectedImprovement(model, best_f=best, sampler=_sampler(128), objective=_generic_objective(w)), X)
    Xb = data.X_baseline.to(DEV)
    _same(qLogNoisyExpectedImprovement(model, X_baseline=Xb, sampler=_sampler(128), objective=lin, prune_baseline=False),
          qLogNoisyExpectedImprovement(model, X_baseline=Xb, sampler=_sampler(128), objective=_generic_objective(w),
                                       prune_baseline=False), X)


@pytest.mark.parametrize("fat", [True, False])
def test_affine_outcome_constraints_inside_the_kernel(fat):
    from botorch_b200.acquisition import qLogExpectedImprovement, qLogNoisyExpectedImprovement
    from botorch_b200.acquisition.monte_carlo import qExpectedImprovement, qPosteriorStandardDeviation

    spec, data, model, X = _setup(S=128)
    med = float(data.train_Y.median())
    cons = [lambda Y: Y[..., 0] - (med + 0.4), lambda Y: -2.0 * Y[..., 0] + 2.0 * (med - 1.5)]   # med - 1.5 <= y <= med + 0.4
    # the reference arm: the same constraints behind an objective the fused route does not recognise (generic torch route)
    from botorch_b200.acquisition.objective import GenericMCObjective

    ident = lambda: GenericMCObjective(lambda Y, X=None: Y[..., 0])  # noqa: E731
    eta = torch.tensor([2e-2, 5e-2])
    best = torch.tensor(med - 0.3, dtype=torch.float64, device=DEV)
    kw = dict(eta=eta, fat=fat)
    f = qLogExpectedImprovement(model, best_f=best, sampler=_sampler(128), constraints=cons, **kw)
    g = qLogExpectedImprovement(model, best_f=best, sampler=_sampler(128), constraints=cons, objective=ident(), **kw)
    assert len(f._fused_constraints()) == 2 and not g._fusable(X)
    a0, b0, e0 = f._fused_constraints()[0]
    assert abs(a0 - 1.0) < 1e-12 and abs(b0 + (med + 0.4)) < 1e-12 and abs(e0 - 2e-2) < 1e-9
    _same(f, g, X)
    Xb = data.X_baseline.to(DEV)

    def nei(**extra):
        # every baseline point violates the first constraint, so `best_f` is the reference's 6-sigma lower bound over 32
        # UNSEEDED uniform points (acquisition/utils.py:181-219): seed the global generator so that both arms draw the same
        torch.manual_seed(11)
        return qLogNoisyExpectedImprovement(model, X_baseline=Xb, sampler=_sampler(128), constraints=cons,
                                            prune_baseline=False, **kw, **extra)

    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")   # "When all training points are infeasible ..."
        _same(nei(), nei(objective=ident()), X)
    if not fat:   # the non-log family multiplies by the plain sigmoid indicator
        _same(qExpectedImprovement(model, best_f=best, sampler=_sampler(128), constraints=cons, eta=eta),
              qExpectedImprovement(model, best_f=best, sampler=_sampler(128), constraints=cons, eta=eta, objective=ident()), X)
        _same(qPosteriorStandardDeviation(model, sampler=_sampler(128), constraints=cons, eta=eta),
              qPosteriorStandardDeviation(model, sampler=_sampler(128), constraints=cons, eta=eta, objective=ident()), X)


def test_non_affine_or_probabilistic_constraints_take_the_generic_route():
    from botorch_b200.acquisition import qLogExpectedImprovement
    from botorch_b200.acquisition._fused import LaunchStats

    spec, data, model, X = _setup(S=64)
    best = torch.tensor(float(data.train_Y.median()), dtype=torch.float64, device=DEV)
    acqf = qLogExpectedImprovement(model, best_f=best, sampler=_sampler(64), constraints=[lambda Y: Y[..., 0] ** 2 - 1.0])
    assert acqf._fused_constraints() is None
    n0 = LaunchStats.launches
    v = acqf(X)
    assert LaunchStats.launches == n0 and torch.isfinite(v).all()
    five = [lambda Y, k=k: Y[..., 0] - k for k in range(5)]   # more than the kernel's four slots
    assert qLogExpectedImprovement(model, best_f=best, sampler=_sampler(64), constraints=five)._fused_constraints() is None


@pytest.mark.parametrize("w", [1.0, -1.3])
def test_mc_mean_utilities_fused_vs_generic_and_reference_formula(w):
    from botorch_b200.acquisition.monte_carlo import (qLowerConfidenceBound, qPosteriorStandardDeviation,
                                                      qUpperConfidenceBound)
    from botorch_b200.acquisition.objective import LinearMCObjective
    from oracle.harness import build_oracle

    spec, data, model, X = _setup("C1", n=None, b=10, S=256)
    lin = None if w == 1.0 else LinearMCObjective(torch.tensor([w], dtype=torch.float64, device=DEV))
    gen = _generic_objective(w)
    for cls, kw in ((qUpperConfidenceBound, dict(beta=2.0)), (qLowerConfidenceBound, dict(beta=0.7)),
                    (qPosteriorStandardDeviation, {})):
        _same(cls(model, sampler=_sampler(256), objective=lin, **kw), cls(model, sampler=_sampler(256), objective=gen, **kw), X)
    # reference formula (monte_carlo.py:896-906) on the ORACLE's posterior samples: mean_S max_q (mu + b' |obj - mu|)
    orc = build_oracle(data)
    Xc = X.cpu()
    mean, cov = orc.gp.posterior_mvn(Xc)
    Lq = torch.linalg.cholesky(cov)
    acqf = qUpperConfidenceBound(model, beta=2.0, sampler=_sampler(256), objective=lin)
    v = acqf(X)
    Z = acqf.sampler.base_samples.reshape(256, spec.q).cpu()
    y = mean.unsqueeze(0) + torch.einsum("bij,sj->sbi", Lq, Z)
    obj = w * y
    mu = obj.mean(dim=0)
    ref = (mu + math.sqrt(2.0 * math.pi / 2) * (obj - mu).abs()).amax(dim=-1).mean(dim=0)
    assert float((v.detach().cpu() - ref).abs().max() / ref.abs().max()) < 1e-9
