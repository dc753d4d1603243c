"""Run-time probe: wherever a REAL `botorch` + `gpytorch` install is importable
(`baseline/_ref` or site-packages -- not in this image: no wheels, no network), the oracle must agree with it to 1e-12 on
the GP boundary (posterior mean / covariance of a `SingleTaskGP` with fixed hyper-parameters) and on qLogEI / qLogNEI values
and gradients with the same Sobol seeds.  Skipped -- visibly -- where the real thing cannot be imported; with it, the
"parity unpinned" caveat of oracle/gp.py is lifted for these cases."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_REF = os.path.join(ROOT, "baseline", "_ref")
if os.path.isdir(_REF) and _REF not in sys.path:
    sys.path.append(_REF)

gpytorch = pytest.importorskip("gpytorch", reason="gpytorch is not installed (no wheel in /opt/wheelhouse, no network)")
botorch = pytest.importorskip("botorch", reason="botorch is not importable without gpytorch / linear_operator")


def _real_model(data):
    from botorch.models import SingleTaskGP
    from botorch.models.transforms.outcome import Standardize
    from gpytorch.kernels import MaternKernel, RBFKernel, ScaleKernel

    d = data.train_X.shape[-1]
    base = (RBFKernel if data.spec.kernel == "rbf" else MaternKernel)(ard_num_dims=d)
    base.lengthscale = data.lengthscale
    covar = base
    if data.spec.outputscale is not None:
        covar = ScaleKernel(base)
        covar.outputscale = data.spec.outputscale
    model = SingleTaskGP(data.train_X, data.train_Y, covar_module=covar, outcome_transform=Standardize(m=1))
    model.likelihood.noise = data.noise
    model.mean_module.constant = 0.0
    return model.eval()


@pytest.mark.parametrize("cfg,n", [("C1", 64), ("C2", 128), ("C3", 256)])
def test_oracle_equals_real_botorch(cfg, n):
    from dataclasses import replace

    from botorch.acquisition.logei import qLogExpectedImprovement, qLogNoisyExpectedImprovement
    from botorch.sampling.normal import SobolQMCNormalSampler

    from botorch_b200.benchmarks import configs
    from oracle.acquisition import value_and_grad
    from oracle.harness import build_oracle

    spec = replace(configs.CONFIGS[cfg], S=64)
    data = configs.make_problem(spec, n=n)
    orc = build_oracle(data)
    model = _real_model(data)
    X = configs.eval_points(data, 6)
    with torch.no_grad():
        post = model.posterior(X)
    m_o, c_o = orc.gp.posterior_mvn(X)
    assert float((post.mean.squeeze(-1) - m_o).abs().max() / m_o.abs().max()) < 1e-12
    cov = post.distribution.covariance_matrix
    assert float((cov - c_o).abs().max() / c_o.abs().max()) < 1e-12
    sampler = SobolQMCNormalSampler(sample_shape=torch.Size([spec.S]), seed=1234)
    if spec.acqf == "qLogEI":
        acqf = qLogExpectedImprovement(model, best_f=data.best_f, sampler=sampler)
    else:
        acqf = qLogNoisyExpectedImprovement(model, X_baseline=data.X_baseline, sampler=sampler, prune_baseline=False)
    Xg = X.clone().requires_grad_(True)
    v = acqf(Xg)
    (g,) = torch.autograd.grad(v.sum(), Xg)
    v_o, g_o = value_and_grad(orc, X)
    assert float(((v.detach() - v_o).abs() / v_o.abs()).max()) < 1e-10
    assert float((g - g_o).abs().max() / g_o.abs().max()) < 1e-8
