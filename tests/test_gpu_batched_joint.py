"""GPU: the batched joint posterior behind everything the fused kernels do not cover (q + r > 32 points per t-batch): values
and gradients against the per-set route and against the oracle, and the generic qLogNEI route (custom objective -> not
fusable) on a t-batch large enough that a per-q-batch Python loop would be noticed."""
import time

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _model(n=300, d=5, kernel="matern52"):
    from dataclasses import replace

    from botorch_b200.benchmarks import configs

    spec = replace(configs.C3, n=n, d=d, kernel=kernel, S=64)
    data = configs.make_problem(spec)
    return data, configs.build_model(data, DEV)


@pytest.mark.parametrize("kernel", ["matern52", "rbf"])
def test_batched_joint_posterior_matches_per_set_route_and_oracle(kernel):
    from oracle.harness import build_oracle

    data, model = _model(kernel=kernel)
    strat = model.prediction_strategy()
    g = torch.Generator().manual_seed(0)
    B, N, d = 7, 45, 5
    X = torch.rand(B, N, d, generator=g, dtype=torch.float64)
    X[0, 1] = X[0, 0]                      # coincident points inside a set (Matern: zero-gradient convention)
    Xg = X.to(DEV).requires_grad_(True)
    mean, covar = strat.batched_joint_posterior(Xg)
    w_m = torch.randn(B, N, generator=g, dtype=torch.float64).to(DEV)
    w_c = torch.randn(B, N, N, generator=g, dtype=torch.float64).to(DEV)
    (gx,) = torch.autograd.grad((mean * w_m).sum() + (covar * w_c).sum(), Xg)
    # per-set route (hand-written backward through the DMMA kernels)
    Xh = X.to(DEV).requires_grad_(True)
    ms, cs = zip(*(strat.joint_posterior_with_grad(x) for x in Xh))
    m2, c2 = torch.stack(ms), torch.stack(cs)
    (gx2,) = torch.autograd.grad((m2 * w_m).sum() + (c2 * w_c).sum(), Xh)
    assert float((mean - m2).abs().max() / m2.abs().max()) < 1e-12
    assert float((covar - c2).abs().max() / c2.abs().max()) < 1e-12
    assert float((gx - gx2).abs().max() / gx2.abs().max()) < 1e-10
    # oracle
    gp = build_oracle(data).gp
    Xo = X.clone().requires_grad_(True)
    m_o, c_o = gp.posterior_mvn(Xo)
    (g_o,) = torch.autograd.grad((m_o * w_m.cpu()).sum() + (c_o * w_c.cpu()).sum(), Xo)
    assert float((mean.detach().cpu() - m_o.detach()).abs().max() / m_o.abs().max()) < 1e-9
    assert float((covar.detach().cpu() - c_o.detach()).abs().max() / c_o.abs().max()) < 1e-9
    assert float((gx.cpu() - g_o).abs().max() / g_o.abs().max()) < 1e-7
    # chunking over the t-batch does not change anything
    m3, c3 = strat.batched_joint_posterior(X.to(DEV), max_rows=2 * N)
    assert torch.equal(m3, mean.detach()) and torch.equal(c3, covar.detach())


def test_generic_qlognei_route_is_batched_over_the_t_batch():
    """A custom (non-affine) objective keeps qLogNEI off the fused kernels; with r = 40 baseline points the joint posterior over
    cat[X_baseline, X] has 44 points per q-batch.  2048 q-batches must go through a handful of launches, not 2048 loops."""
    from botorch_b200.acquisition import qLogNoisyExpectedImprovement
    from botorch_b200.acquisition.objective import GenericMCObjective
    from botorch_b200.sampling import SobolQMCNormalSampler
    from oracle.acquisition import OracleQLogNEI

    data, model = _model(n=256, d=5)
    g = torch.Generator().manual_seed(1)
    Xb = torch.rand(40, 5, generator=g, dtype=torch.float64)
    obj = GenericMCObjective(lambda samples, X=None: samples.squeeze(-1))   # identity, but opaque to the fusability gate
    acqf = qLogNoisyExpectedImprovement(model, X_baseline=Xb.to(DEV), prune_baseline=False, objective=obj,
                                        sampler=SobolQMCNormalSampler(torch.Size([32]), seed=3))
    Xq = torch.rand(2048, 4, 5, generator=g, dtype=torch.float64).to(DEV)
    acqf(Xq[:8])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    Xg = Xq.clone().requires_grad_(True)
    v = acqf(Xg)
    (gr,) = torch.autograd.grad(v.sum(), Xg)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert v.shape == (2048,) and torch.isfinite(v).all() and torch.isfinite(gr).all()
    assert dt < 5.0, f"generic route took {dt:.1f} s for 2048 q-batches"
    # parity of a slice with the oracle (identity objective)
    from oracle.harness import build_oracle

    gp = build_oracle(data).gp
    orc = OracleQLogNEI(gp, Xb, 32, 3)
    v_o = orc(Xq[:16].cpu())
    assert float(((v[:16].detach().cpu() - v_o).abs() / v_o.abs()).max()) < 1e-8
