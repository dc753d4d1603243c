"""Config 5 (shrunk n): ModelListGP with 4 outputs + qLogEHVI-style MC objective.  The CUDA path (4 CUDA posteriors,
fused CUDA log-areas kernel) against the oracle (4 oracle GPs + log-areas restatement pinned to the compiled reference)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _problem(n=192, d=6, m=4, q=3, S=64, nc=12, b=5):
    from botorch_b200.acquisition.multi_objective import qLogExpectedHypervolumeImprovement
    from botorch_b200.models import MaternKernel, ModelListGP, ScaleKernel, SingleTaskGP
    from botorch_b200.sampling import SobolQMCNormalSampler
    from oracle.gp import OracleGP
    from oracle.mo import OracleQLogEHVI

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    X = torch.rand(n, d, generator=g, dtype=torch.float64)
    models, gps = [], []
    for k in range(m):
        Y = torch.sin((k + 2) * X.sum(-1, keepdim=True)) + 0.1 * (k + 1) * X[:, k % d: k % d + 1] \
            + 0.02 * torch.randn(n, 1, generator=g, dtype=torch.float64)
        ls = 0.4 + 0.5 * torch.rand(d, generator=g, dtype=torch.float64)
        mod = SingleTaskGP(X.to(dev), Y.to(dev), covar_module=ScaleKernel(MaternKernel(ard_num_dims=d, lengthscale=ls), outputscale=1.0 + 0.3 * k))
        mod.likelihood.noise = 1e-3
        models.append(mod.to(dev))
        gps.append(OracleGP(X, Y, ls, torch.tensor(1e-3, dtype=torch.float64), kernel="matern52", outputscale=1.0 + 0.3 * k))
    lo = torch.randn(nc, m, generator=g, dtype=torch.float64) * 0.5 - 0.5
    hi = lo + torch.rand(nc, m, generator=g, dtype=torch.float64) + 0.1
    hi[-1] = float("inf")
    acqf = qLogExpectedHypervolumeImprovement(ModelListGP(*models), cell_bounds=(lo.to(dev), hi.to(dev)),
                                              sampler=SobolQMCNormalSampler(torch.Size([S]), seed=1234))
    orc = OracleQLogEHVI(gps, lo, hi, S, 1234)
    Xq = torch.rand(b, q, d, generator=g, dtype=torch.float64)
    return acqf, orc, Xq, dev


def test_qlogehvi_value_and_grad_match_oracle():
    acqf, orc, Xq, dev = _problem()
    Xo = Xq.clone().requires_grad_(True)
    v_o = orc(Xo)
    (g_o,) = torch.autograd.grad(v_o.sum(), Xo)
    Xg = Xq.to(dev).requires_grad_(True)
    v = acqf(Xg)
    (g,) = torch.autograd.grad(v.sum(), Xg)
    assert v.shape == (Xq.shape[0],)
    assert float(((v.detach().cpu() - v_o.detach()).abs() / v_o.detach().abs()).max()) < 1e-9
    assert float((g.cpu() - g_o).abs().max() / g_o.abs().max()) < 1e-7


def test_model_list_posterior_shapes():
    acqf, orc, Xq, dev = _problem(b=3)
    post = acqf.model.posterior(Xq.to(dev))
    assert post.mean.shape == (3, 3, 4) and post.variance.shape == (3, 3, 4)
    assert acqf.model.num_outputs == 4
    for k, gp in enumerate(orc.gps):
        m_o, c_o = gp.posterior_mvn(Xq)
        assert float((post.mean[..., k].cpu() - m_o).abs().max() / m_o.abs().max()) < 1e-9
