"""Edge shapes of the fused path against the oracle: minimum / maximum q, r, d, odd n (padding of the contraction
dimension), a single MC sample, ragged and empty batches, both contraction modes."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _problem(n, d, seed, ls_factor=1.0):
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(n, d, generator=g, dtype=torch.float64)
    Y = torch.sin(3.0 * X.sum(-1, keepdim=True) / d ** 0.5) + 0.1 * torch.randn(n, 1, generator=g, dtype=torch.float64)
    if n == 1:
        Y = Y + 0.5
    # short lengthscales keep the q x q conditional covariances well away from singular, so that no jitter decision
    # (a rounding-level coin flip in the reference as well) enters the comparison
    ls = (0.2 + 0.2 * torch.rand(d, generator=g, dtype=torch.float64)) * d ** 0.5 * ls_factor
    return X, Y, ls, g


def _pair(n, d, kernel, seed, contraction, ls_factor=1.0):
    from botorch_b200 import settings
    from botorch_b200.models import MaternKernel, RBFKernel, SingleTaskGP
    from botorch_b200.models.transforms import Standardize
    from oracle.gp import OracleGP

    X, Y, ls, g = _problem(n, d, seed, ls_factor)
    cov = (RBFKernel if kernel == "rbf" else MaternKernel)(ard_num_dims=d, lengthscale=ls)
    settings.contraction.set(contraction)
    model = SingleTaskGP(X.to(DEV), Y.to(DEV), covar_module=cov, outcome_transform=Standardize(m=1)).to(DEV)
    model.likelihood.noise = 1e-2
    gp = OracleGP(X, Y, ls, torch.tensor(1e-2, dtype=torch.float64), kernel=kernel)
    return model, gp, X, Y, g


# (n, d, q, r, S, b, kernel)
SHAPES = [
    (1, 1, 1, 0, 1, 1, "rbf"),          # everything minimal
    (17, 3, 1, 0, 8, 5, "matern52"),    # q = 1, odd n (np padded to 32), b not a multiple of the 4-batch warp tile
    (33, 64, 32, 0, 16, 3, "rbf"),      # maximum d and q
    (50, 5, 5, 64, 32, 2, "matern52"),  # maximum r
    (100, 2, 9, 17, 33, 7, "rbf"),      # odd q / r / S (26 jointly sampled points in the unit square: short lengthscale below)
    (130, 6, 17, 40, 24, 4, "matern52"),
    (200, 8, 32, 64, 16, 2, "matern52"),  # maximum q AND maximum r (largest register / shared-memory variants)
    (64, 4, 4, 8, 1000, 3, "rbf"),        # S not a multiple of the 128-thread sample pass
]


@pytest.mark.parametrize("contraction", ["dmma", "int8"])
@pytest.mark.parametrize("n,d,q,r,S,b,kernel", SHAPES)
def test_edge_shapes_value_and_grad(n, d, q, r, S, b, kernel, contraction):
    from botorch_b200 import settings
    from botorch_b200.acquisition import qLogExpectedImprovement, qLogNoisyExpectedImprovement
    from botorch_b200.sampling import SobolQMCNormalSampler
    from oracle.acquisition import OracleQLogEI, OracleQLogNEI, value_and_grad

    try:
        model, gp, X, Y, g = _pair(n, d, kernel, 11 * n + q, contraction, ls_factor=0.3 if d == 2 else 1.0)
        sampler = SobolQMCNormalSampler(torch.Size([S]), seed=3)
        if r == 0:
            best = Y.max()
            acqf = qLogExpectedImprovement(model, best_f=best.to(DEV), sampler=sampler)
            orc = OracleQLogEI(gp, best, S, 3)
        else:
            Xb = torch.rand(r, d, generator=g, dtype=torch.float64)
            acqf = qLogNoisyExpectedImprovement(model, X_baseline=Xb.to(DEV), prune_baseline=False, sampler=sampler)
            orc = OracleQLogNEI(gp, Xb, S, 3)
        Xq = torch.rand(b, q, d, generator=g, dtype=torch.float64)
        v_o, g_o = value_and_grad(orc, Xq)
        Xg = Xq.to(DEV).requires_grad_(True)
        v = acqf(Xg)
        assert v.shape == (b,)
        (gr,) = torch.autograd.grad(v.sum(), Xg)
        # one bar for both contraction modes (round 2: the int8 mode picks its slice count per model): 1e-8 / 1e-6 at these
        # tiny n, where small pivots of the q x q conditional covariances amplify any rounding of the posterior blocks
        vtol, gtol = 1e-8, 1e-6
        assert float(((v.detach().cpu() - v_o).abs() / v_o.abs().clamp_min(1e-300)).max()) < vtol
        assert float((gr.cpu() - g_o).abs().max() / g_o.abs().max().clamp_min(1e-300)) < gtol
        # empty t-batch: shape-only result, no launch
        empty = acqf(torch.empty(0, q, d, device=DEV, dtype=torch.float64))
        assert empty.shape == (0,)
    finally:
        settings.contraction.set("dmma")


def test_baseline_beyond_kernel_limit_takes_generic_route():
    """r > MCACQ_MAX_R is outside the fused kernels: the host falls back to the generic sample-reducing route, whose joint
    posterior over cat[X, X_baseline] (and its gradient) is assembled from the covariance / contraction kernels."""
    from botorch_b200 import _lib
    from botorch_b200.acquisition import qLogNoisyExpectedImprovement
    from botorch_b200.sampling import SobolQMCNormalSampler
    from oracle.acquisition import OracleQLogNEI, value_and_grad

    model, gp, X, Y, g = _pair(90, 3, "rbf", 5, "dmma")
    r = _lib.MAX_R + 6
    Xb = torch.rand(r, 3, generator=g, dtype=torch.float64)
    acqf = qLogNoisyExpectedImprovement(model, X_baseline=Xb.to(DEV), prune_baseline=False,
                                        sampler=SobolQMCNormalSampler(torch.Size([16]), seed=3))
    Xq = torch.rand(3, 2, 3, generator=g, dtype=torch.float64)
    v_o, g_o = value_and_grad(OracleQLogNEI(gp, Xb, 16, 3), Xq)
    Xg = Xq.to(DEV).requires_grad_(True)
    v = acqf(Xg)
    (gr,) = torch.autograd.grad(v.sum(), Xg)
    assert float(((v.detach().cpu() - v_o).abs() / v_o.abs()).max()) < 1e-7
    assert float((gr.cpu() - g_o).abs().max() / g_o.abs().max()) < 1e-5
    # the joint posterior behind it (q + r = 72 points per t-batch) against the oracle, values and gradient
    Xj = torch.cat([Xq, Xb.expand(3, r, 3)], dim=-2)
    m_o, c_o = gp.posterior_mvn(Xj)
    Xjg = Xj.to(DEV).requires_grad_(True)
    post = model.posterior(Xjg)
    assert float((post.mean.squeeze(-1).detach().cpu() - m_o).abs().max() / m_o.abs().max()) < 1e-9
    assert float((post.distribution.covariance_matrix.detach().cpu() - c_o).abs().max() / c_o.abs().max()) < 1e-9
    w = torch.randn(c_o.shape, generator=g, dtype=torch.float64)
    Xjo = Xj.clone().requires_grad_(True)
    m2, c2 = gp.posterior_mvn(Xjo)
    (g_ref,) = torch.autograd.grad((c2 * w).sum() + m2.sum(), Xjo)
    (g_dev,) = torch.autograd.grad((post.distribution.covariance_matrix * w.to(DEV)).sum() + post.mean.sum(), Xjg)
    assert float((g_dev.cpu() - g_ref).abs().max() / g_ref.abs().max()) < 1e-8


def test_int8_guard_falls_back_on_ill_conditioned_models():
    """The int8 slices are fixed-point per row: on a nearly singular train covariance the probe at strategy build time
    must detect the lost digits and switch the model to the FP64 contraction (with a NumericalWarning)."""
    import warnings

    from botorch_b200 import settings
    from botorch_b200.exceptions import NumericalWarning
    from botorch_b200.models import RBFKernel, SingleTaskGP

    g = torch.Generator().manual_seed(0)
    X = torch.rand(300, 1, generator=g, dtype=torch.float64)
    Y = torch.sin(6 * X)
    try:
        settings.contraction.set("int8")
        smooth = SingleTaskGP(X.to(DEV), Y.to(DEV), covar_module=RBFKernel(ard_num_dims=1, lengthscale=torch.tensor([1.5]))).to(DEV)
        smooth.likelihood.noise = 1e-6
        with warnings.catch_warnings(record=True) as ws:
            warnings.simplefilter("always")
            strat = smooth.prediction_strategy()
        assert strat.contraction == "dmma" and strat.int8_var_ratio_limit is None
        assert any(issubclass(w.category, NumericalWarning) and "int8" in str(w.message) for w in ws)
        rough = SingleTaskGP(X.to(DEV), Y.to(DEV), covar_module=RBFKernel(ard_num_dims=1, lengthscale=torch.tensor([0.02]))).to(DEV)
        rough.likelihood.noise = 1e-2
        strat = rough.prediction_strategy()
        assert strat.contraction == "int8" and strat.int8_probe_error <= strat.INT8_PROBE_TOL
    finally:
        settings.contraction.set("dmma")


def test_chunk_boundary_is_invisible():
    """b * q beyond the 2^17-row workspace chunk: the host splits the t-batch; results must equal the pieces evaluated alone
    (bitwise) and the oracle on a sample of rows."""
    from botorch_b200.acquisition import qLogNoisyExpectedImprovement
    from botorch_b200.sampling import SobolQMCNormalSampler
    from oracle.acquisition import OracleQLogNEI, value_and_grad

    model, gp, X, Y, g = _pair(64, 3, "matern52", 21, "dmma")
    Xb = torch.rand(5, 3, generator=g, dtype=torch.float64)
    acqf = qLogNoisyExpectedImprovement(model, X_baseline=Xb.to(DEV), prune_baseline=False,
                                        sampler=SobolQMCNormalSampler(torch.Size([32]), seed=3))
    b, q = 33000, 4  # 132000 rows: one full chunk of 32768 q-batches + a ragged tail of 232
    Xq = torch.rand(b, q, 3, generator=g, dtype=torch.float64).to(DEV)
    Xg = Xq.clone().requires_grad_(True)
    v = acqf(Xg)
    (gr,) = torch.autograd.grad(v.sum(), Xg)
    for lo, hi in ((0, 32768), (32768, b)):
        Xp = Xq[lo:hi].clone().requires_grad_(True)
        vp = acqf(Xp)
        (gp_,) = torch.autograd.grad(vp.sum(), Xp)
        assert torch.equal(vp.detach(), v.detach()[lo:hi]) and torch.equal(gp_, gr[lo:hi])
    idx = torch.tensor([0, 1, 32767, 32768, 32769, b - 1])
    v_o, g_o = value_and_grad(OracleQLogNEI(gp, Xb, 32, 3), Xq[idx].cpu())
    assert float(((v.detach()[idx].cpu() - v_o).abs() / v_o.abs()).max()) < 1e-8
    assert float((gr[idx].cpu() - g_o).abs().max() / g_o.abs().max()) < 1e-6


@pytest.mark.parametrize("contraction", ["dmma", "int8"])
def test_wide_and_narrow_sample_reduce_are_bit_identical(contraction, monkeypatch):
    """Optimiser rounds (b < 296 q-batches) run the sample/reduce kernels with 4 x 128 threads per q-batch, sweeps with
    128: the per-thread fold order is the same by construction, so values and gradients must agree bit for bit
    (this is what keeps a t-batch's result independent of how it is chunked or sharded)."""
    from botorch_b200 import settings
    from botorch_b200.acquisition import qExpectedImprovement, qLogExpectedImprovement, qLogNoisyExpectedImprovement
    from botorch_b200.sampling import SobolQMCNormalSampler

    try:
        model, gp, X, Y, g = _pair(90, 4, "matern52", 3, contraction)
        Xb = torch.rand(11, 4, generator=g, dtype=torch.float64).to(DEV)
        mk = lambda S: SobolQMCNormalSampler(torch.Size([S]), seed=2)
        cases = [(qLogNoisyExpectedImprovement(model, X_baseline=Xb, prune_baseline=False, sampler=mk(1024)), 8),
                 (qLogExpectedImprovement(model, best_f=float(Y.max()), sampler=mk(200)), 3),
                 (qLogExpectedImprovement(model, best_f=float(Y.max()), sampler=mk(512), fat=False), 12),
                 (qExpectedImprovement(model, best_f=float(Y.median()), sampler=mk(300)), 5)]
        for acqf, q in cases:
            Xq = torch.rand(9, q, 4, generator=g, dtype=torch.float64).to(DEV)
            out = {}
            for wide in ("0", "1"):
                monkeypatch.setenv("MCACQ_SR_WIDE", wide)
                Xg = Xq.clone().requires_grad_(True)
                v = acqf(Xg)
                (gr,) = torch.autograd.grad(v.sum(), Xg)
                out[wide] = (v.detach().clone(), gr.clone())
            assert torch.equal(out["0"][0], out["1"][0]) and torch.equal(out["0"][1], out["1"][1])
            assert torch.isfinite(out["1"][0]).all()
    finally:
        settings.contraction.set("dmma")
