"""Outcome constraints on the generic sample-reducing route: a 2-output ModelListGP (objective =
output 0, constraint on output 1), CUDA posteriors + host constraint weighting, against the oracle."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
CONS = [lambda Y: Y[..., 1] - 0.15]


def _setup(n=80, d=3, seed=4):
    from botorch_b200.models import MaternKernel, ModelListGP, SingleTaskGP
    from oracle.gp import OracleGP

    g = torch.Generator().manual_seed(seed)
    X = torch.rand(n, d, generator=g, dtype=torch.float64)
    models, gps, Ys = [], [], []
    for k in range(2):
        Y = torch.sin((2 + k) * X.sum(-1, keepdim=True)) + 0.3 * k * X[:, :1] + 0.03 * torch.randn(n, 1, generator=g, dtype=torch.float64)
        ls = 0.3 + 0.3 * torch.rand(d, generator=g, dtype=torch.float64)
        mod = SingleTaskGP(X.to(DEV), Y.to(DEV), covar_module=MaternKernel(ard_num_dims=d, lengthscale=ls)).to(DEV)
        mod.likelihood.noise = 5e-3
        models.append(mod)
        gps.append(OracleGP(X, Y, ls, torch.tensor(5e-3, dtype=torch.float64), kernel="matern52"))
        Ys.append(Y)
    return ModelListGP(*models), gps, X, Ys, g


def _objective():
    from botorch_b200.acquisition.objective import GenericMCObjective

    return GenericMCObjective(lambda Y, X=None: Y[..., 0])


@pytest.mark.parametrize("log", [True, False])
def test_constrained_qei_matches_oracle(log):
    from botorch_b200.acquisition import qLogExpectedImprovement
    from botorch_b200.acquisition.monte_carlo import qExpectedImprovement
    from botorch_b200.sampling import SobolQMCNormalSampler
    from oracle.constraints import OracleConstrainedQEI

    model, gps, X, Ys, g = _setup()
    feas = Ys[1].squeeze(-1) <= 0.15
    best = Ys[0].squeeze(-1)[feas].max()
    S, eta = 64, 2e-2
    cls = qLogExpectedImprovement if log else qExpectedImprovement
    acqf = cls(model, best_f=best.to(DEV), sampler=SobolQMCNormalSampler(torch.Size([S]), seed=9), objective=_objective(),
               constraints=CONS, eta=eta)
    orc = OracleConstrainedQEI(gps, CONS, best, S, 9, log=log, eta=eta)
    Xq = torch.rand(6, 3, 3, generator=g, dtype=torch.float64)
    Xo = Xq.clone().requires_grad_(True)
    v_o = orc(Xo)
    (g_o,) = torch.autograd.grad(v_o.sum(), Xo)
    Xg = Xq.to(DEV).requires_grad_(True)
    v = acqf(Xg)
    (gr,) = torch.autograd.grad(v.sum(), Xg)
    assert float(((v.detach().cpu() - v_o.detach()).abs() / v_o.detach().abs().clamp_min(1e-12)).max()) < 1e-9
    assert float((gr.cpu() - g_o).abs().max() / g_o.abs().max()) < 1e-7
    # the constraint matters: the unconstrained value differs
    free = cls(model, best_f=best.to(DEV), sampler=SobolQMCNormalSampler(torch.Size([S]), seed=9), objective=_objective())
    assert not torch.allclose(free(Xq.to(DEV)), v.detach())


def test_constrained_qlognei_best_feasible_and_pruning():
    from botorch_b200.acquisition import qLogNoisyExpectedImprovement
    from botorch_b200.acquisition.utils import compute_best_feasible_objective, prune_inferior_points
    from botorch_b200.exceptions import BotorchWarning
    from botorch_b200.sampling import SobolQMCNormalSampler

    model, gps, X, Ys, g = _setup()
    Xb = X[:12].to(DEV)
    acqf = qLogNoisyExpectedImprovement(model, X_baseline=Xb, objective=_objective(), constraints=CONS, eta=2e-2,
                                        prune_baseline=True, sampler=SobolQMCNormalSampler(torch.Size([32]), seed=2))
    assert 1 <= acqf.X_baseline.shape[0] <= 12
    Xq = torch.rand(4, 2, 3, generator=g, dtype=torch.float64).to(DEV).requires_grad_(True)
    v = acqf(Xq)
    (gr,) = torch.autograd.grad(v.sum(), Xq)
    assert v.shape == (4,) and torch.isfinite(v).all() and torch.isfinite(gr).all()
    # best feasible objective: feasible entries only; all-infeasible falls back to a lower bound (with a warning)
    samples = torch.tensor([[[1.0, 1.0], [3.0, 0.0], [2.0, -1.0]]], dtype=torch.float64, device=DEV)  # 1 x 3 x 2
    obj = samples[..., 0]
    assert float(compute_best_feasible_objective(samples, obj, constraints=[lambda Y: Y[..., 1]])) == 3.0
    assert float(compute_best_feasible_objective(samples, obj, constraints=[lambda Y: Y[..., 1] + 0.5])) == 2.0
    none_ok = [lambda Y: Y[..., 1] + 10.0]
    assert float(compute_best_feasible_objective(samples, obj, constraints=none_ok, infeasible_obj=torch.tensor(-7.0))) == -7.0
    with pytest.warns(BotorchWarning):
        lb = compute_best_feasible_objective(samples, obj, constraints=none_ok, model=model, objective=_objective(), X_baseline=Xb)
    assert float(lb) < float(Ys[0].min())
    with pytest.raises(ValueError):
        compute_best_feasible_objective(samples, obj, constraints=none_ok)
    # pruning with constraints keeps only points that can be the best FEASIBLE one
    kept = prune_inferior_points(model, X[:20].to(DEV), objective=_objective(), constraints=CONS)
    assert 1 <= kept.shape[0] <= 20
