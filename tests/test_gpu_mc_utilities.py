"""N3: qEI / qNEI / qSimpleRegret / qProbabilityOfImprovement on the fused kernel against the oracle restatements
(values 1e-9 relative to the value scale, gradients 1e-7), and fused vs generic route."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(cfg, b=24, **over):
    from dataclasses import replace

    from botorch_b200.benchmarks import configs
    from oracle.harness import build_oracle

    spec = replace(configs.CONFIGS[cfg], **over)
    data = configs.make_problem(spec)
    dev = torch.device("cuda:0")
    model = configs.build_model(data, dev)
    return data, model, build_oracle(data), configs.eval_points(data, b), dev


def _check(acqf, oracle_fn, X, dev, vtol=1e-9, gtol=1e-7):
    Xo = X.clone().requires_grad_(True)
    v_o = oracle_fn(Xo)
    (g_o,) = torch.autograd.grad(v_o.sum(), Xo)
    Xg = X.to(dev).requires_grad_(True)
    v = acqf(Xg)
    (g,) = torch.autograd.grad(v.sum(), Xg)
    scale = v_o.detach().abs().max().clamp_min(1e-300)
    assert float((v.detach().cpu() - v_o.detach()).abs().max() / scale) < vtol
    assert float((g.cpu() - g_o).abs().max() / g_o.abs().max().clamp_min(1e-300)) < gtol


def test_qei_qsr_qpi_match_oracle():
    from botorch_b200.acquisition import qExpectedImprovement, qProbabilityOfImprovement, qSimpleRegret
    from botorch_b200.sampling import SobolQMCNormalSampler
    from oracle import acquisition as oa

    data, model, orc, X, dev = _setup("C1", S=256)
    # an incumbent inside the sample range so that relu / sigmoid are active for a good share of the samples
    best = torch.tensor(float(data.train_Y.median()), dtype=torch.float64)
    orc.best_f = best
    mk = lambda: SobolQMCNormalSampler(torch.Size([256]), seed=1234)
    _check(qExpectedImprovement(model, best_f=best.to(dev), sampler=mk()), lambda x: oa.oracle_qei(orc, x), X, dev)
    _check(qSimpleRegret(model, sampler=mk()), lambda x: oa.oracle_qsr(orc, x), X, dev)
    _check(qProbabilityOfImprovement(model, best_f=best.to(dev), sampler=mk(), tau=1e-2),
           lambda x: oa.oracle_qpi(orc, x, tau=1e-2), X, dev, vtol=1e-8, gtol=1e-6)


def test_qnei_matches_oracle_and_generic_route():
    from botorch_b200.acquisition import GenericMCObjective, qNoisyExpectedImprovement
    from botorch_b200.sampling import SobolQMCNormalSampler
    from oracle import acquisition as oa

    data, model, orc, X, dev = _setup("C2", n=256, S=128, r=8)
    # a weak incumbent set (the 8 WORST training points) so that the noisy improvement is non-zero
    from oracle.acquisition import OracleQLogNEI

    data.X_baseline = data.train_X[(-data.train_Y.squeeze(-1)).topk(8).indices].clone()
    orc = OracleQLogNEI(orc.gp, data.X_baseline, 128, 1234)
    mk = lambda: SobolQMCNormalSampler(torch.Size([128]), seed=1234)
    acqf = qNoisyExpectedImprovement(model, X_baseline=data.X_baseline.to(dev), sampler=mk(), prune_baseline=False)
    _check(acqf, lambda x: oa.oracle_qnei(orc, x), X, dev)
    generic = qNoisyExpectedImprovement(model, X_baseline=data.X_baseline.to(dev), sampler=mk(), prune_baseline=False,
                                        objective=GenericMCObjective(lambda s, X=None: s.squeeze(-1)))
    Xa, Xb = X.to(dev).requires_grad_(True), X.to(dev).requires_grad_(True)
    va, vb = acqf(Xa), generic(Xb)
    (ga,), (gb,) = torch.autograd.grad(va.sum(), Xa), torch.autograd.grad(vb.sum(), Xb)
    assert float(vb.detach().abs().max()) > 0
    assert float((va - vb).detach().abs().max() / vb.detach().abs().max()) < 1e-9
    assert float((ga - gb).abs().max() / gb.abs().max()) < 1e-7


def test_qucb_qlcb_qpstd_match_oracle():
    """Utilities that depend on the MC mean over all samples: generic route on CUDA posteriors against the oracle."""
    from botorch_b200.acquisition import qLowerConfidenceBound, qPosteriorStandardDeviation, qUpperConfidenceBound
    from botorch_b200.sampling import SobolQMCNormalSampler
    from oracle import acquisition as oa

    data, model, orc, X, dev = _setup("C1", S=256)
    mk = lambda: SobolQMCNormalSampler(torch.Size([256]), seed=1234)
    _check(qUpperConfidenceBound(model, beta=2.0, sampler=mk()), lambda x: oa.oracle_qucb(orc, x, 2.0), X, dev)
    _check(qLowerConfidenceBound(model, beta=2.0, sampler=mk()), lambda x: oa.oracle_qucb(orc, x, 2.0, lower=True), X, dev)
    _check(qPosteriorStandardDeviation(model, sampler=mk()), lambda x: oa.oracle_qpstd(orc, x), X, dev)
