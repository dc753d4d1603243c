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


@pytest.mark.parametrize("q,m,nc", [(1, 2, 5), (2, 3, 7), (3, 4, 12), (4, 4, 32), (5, 2, 9), (6, 3, 4)])
def test_fused_log_hvi_kernel_equals_the_per_subset_size_route(q, m, nc):
    """csrc/log_hvi.cu (one launch: subsets as bit masks, streaming log-sum-exps) against the route that mirrors the reference
    step by step (per-subset-size `fused_log_areas` launches + safe_math reductions): per-sample values and gradients w.r.t.
    the objective samples, incl. a cell with an infinite upper bound and samples far below / above the cells."""
    from botorch_b200.acquisition.multi_objective.fused_log_areas import fused_log_hvi
    from botorch_b200.acquisition.multi_objective.logei import qLogExpectedHypervolumeImprovement

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(10 * q + m)
    lo = torch.randn(nc, m, generator=g, dtype=torch.float64) * 0.5 - 0.5
    hi = lo + torch.rand(nc, m, generator=g, dtype=torch.float64) + 0.1
    hi[-1] = float("inf")
    B = 257
    obj = torch.randn(B, q, m, generator=g, dtype=torch.float64)
    obj[0] = -40.0          # far below every cell: log_fatplus deep in its tail
    obj[1] = 40.0           # far above
    obj[2, 0] = obj[2, -1]  # tie between two points
    acqf = qLogExpectedHypervolumeImprovement.__new__(qLogExpectedHypervolumeImprovement)
    torch.nn.Module.__init__(acqf)
    acqf.register_buffer("cell_lower_bounds", lo.to(dev))
    acqf.register_buffer("cell_upper_bounds", hi.to(dev))
    acqf.tau_relu, acqf.tau_max, acqf.q_out, acqf.q_subset_indices = 1e-6, 1e-2, -1, {}
    from botorch_b200 import settings

    o1 = obj.to(dev).requires_grad_(True)
    v1 = fused_log_hvi(o1, lo.to(dev), hi.to(dev), 1e-6, 1e-2)
    w = torch.randn(B, generator=g, dtype=torch.float64).to(dev)
    (g1,) = torch.autograd.grad((v1 * w).sum(), o1)
    o2 = obj.to(dev).requires_grad_(True)
    with settings.fused_log_hvi(False):
        # the reference-shaped route reduces over dim 0 at the end (logmeanexp over S = 1 is the identity)
        v2 = acqf._compute_log_qehvi(o2.unsqueeze(0).transpose(0, 1).reshape(1, B, q, m))
    (g2,) = torch.autograd.grad((v2 * w).sum(), o2)
    assert torch.isfinite(v1).all() and torch.isfinite(g1).all()
    assert float(((v1 - v2).abs() / v2.abs().clamp_min(1e-12)).max()) < 1e-11
    assert float((g1 - g2).abs().max() / g2.abs().max()) < 1e-10
