"""Drop-in boundary semantics on the GPU against the oracle: Normalize input transform on a non-unit box, fixed
observation noise (`train_Yvar`), `fat=False` (log_softplus + smooth_amax), iid base samples, pending points, jitter
signalling, and error conventions."""
import warnings

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _data(n=96, d=4, seed=0, box=(-3.0, 5.0)):
    g = torch.Generator().manual_seed(seed)
    lo, hi = box
    X = lo + (hi - lo) * torch.rand(n, d, generator=g, dtype=torch.float64)
    Y = torch.sin(X.sum(-1, keepdim=True) * 0.7) + 0.3 * X[:, :1] + 0.05 * torch.randn(n, 1, generator=g, dtype=torch.float64)
    ls = 0.25 + 0.3 * torch.rand(d, generator=g, dtype=torch.float64)
    bounds = torch.tensor([[lo] * d, [hi] * d], dtype=torch.float64)
    return X, Y, ls, bounds, g


def test_normalize_transform_and_fixed_noise_parity():
    from botorch_b200.acquisition import qLogNoisyExpectedImprovement
    from botorch_b200.models import RBFKernel, SingleTaskGP
    from botorch_b200.models.transforms import Normalize
    from botorch_b200.sampling import SobolQMCNormalSampler
    from oracle.acquisition import OracleQLogNEI, value_and_grad
    from oracle.gp import OracleGP

    X, Y, ls, bounds, g = _data()
    Yvar = 1e-3 + 4e-3 * torch.rand(X.shape[0], 1, generator=g, dtype=torch.float64)
    model = SingleTaskGP(X.to(DEV), Y.to(DEV), train_Yvar=Yvar.to(DEV), covar_module=RBFKernel(ard_num_dims=4, lengthscale=ls),
                         input_transform=Normalize(d=4, bounds=bounds.to(DEV))).to(DEV)
    s = float(Y.std())
    gp = OracleGP(X, Y, ls, (Yvar / s**2).squeeze(-1), kernel="rbf", norm_offset=bounds[0], norm_coef=bounds[1] - bounds[0])
    Xb = X[Y.squeeze(-1).topk(6).indices]
    acqf = qLogNoisyExpectedImprovement(model, X_baseline=Xb.to(DEV), prune_baseline=False,
                                        sampler=SobolQMCNormalSampler(torch.Size([128]), seed=7))
    orc = OracleQLogNEI(gp, Xb, 128, 7)
    Xq = bounds[0] + (bounds[1] - bounds[0]) * torch.rand(9, 3, 4, generator=g, dtype=torch.float64)
    post = model.posterior(Xq.to(DEV))
    m_o, c_o = gp.posterior_mvn(Xq)
    assert float((post.mean.squeeze(-1).cpu() - m_o).abs().max() / m_o.abs().max()) < 1e-9
    assert float((post.distribution.covariance_matrix.cpu() - c_o).abs().max() / c_o.abs().max()) < 1e-9
    v_o, g_o = value_and_grad(orc, Xq)
    Xg = Xq.to(DEV).requires_grad_(True)
    v = acqf(Xg)
    (gr,) = torch.autograd.grad(v.sum(), Xg)
    assert float(((v.detach().cpu() - v_o).abs() / v_o.abs()).max()) < 1e-9
    assert float((gr.cpu() - g_o).abs().max() / g_o.abs().max()) < 1e-7


def test_non_fat_variant_and_pending_points():
    from botorch_b200.acquisition import qLogExpectedImprovement
    from botorch_b200.models import MaternKernel, ScaleKernel, SingleTaskGP
    from botorch_b200.sampling import SobolQMCNormalSampler
    from oracle.acquisition import OracleQLogEI, value_and_grad
    from oracle.gp import OracleGP

    X, Y, ls, bounds, g = _data(box=(0.0, 1.0), seed=3)
    model = SingleTaskGP(X.to(DEV), Y.to(DEV), covar_module=ScaleKernel(MaternKernel(ard_num_dims=4, lengthscale=ls), 1.4)).to(DEV)
    model.likelihood.noise = 2e-3
    gp = OracleGP(X, Y, ls, torch.tensor(2e-3, dtype=torch.float64), kernel="matern52", outputscale=1.4)
    best = Y.max()
    Xq = torch.rand(7, 2, 4, generator=g, dtype=torch.float64)
    # fat=False: log_softplus + smooth_amax (safe_math.py:228-278)
    acqf = qLogExpectedImprovement(model, best_f=best.to(DEV), fat=False, tau_relu=1e-3,
                                   sampler=SobolQMCNormalSampler(torch.Size([64]), seed=5))
    orc = OracleQLogEI(gp, best, 64, 5, tau_relu=1e-3, fat=False)
    v_o, g_o = value_and_grad(orc, Xq)
    Xg = Xq.to(DEV).requires_grad_(True)
    v = acqf(Xg)
    (gr,) = torch.autograd.grad(v.sum(), Xg)
    assert float(((v.detach().cpu() - v_o).abs() / v_o.abs()).max()) < 1e-9
    assert float((gr.cpu() - g_o).abs().max() / g_o.abs().max()) < 1e-7
    # pending points are appended along q (utils/transforms.py:378-403): acqf(X) with X_pending == acqf(cat[X, P])
    fat = qLogExpectedImprovement(model, best_f=best.to(DEV), sampler=SobolQMCNormalSampler(torch.Size([64]), seed=5))
    P = torch.rand(1, 4, generator=g, dtype=torch.float64).to(DEV)
    with torch.no_grad():
        joint = fat(torch.cat([Xq.to(DEV), P.expand(7, 1, 4)], dim=-2))
        fat.set_X_pending(P)
        pend = fat(Xq.to(DEV))
    assert torch.equal(joint, pend)


def test_iid_sampler_and_jitter_signalling():
    from botorch_b200.acquisition import qLogExpectedImprovement
    from botorch_b200.exceptions import NumericalWarning
    from botorch_b200.models import RBFKernel, SingleTaskGP
    from botorch_b200.sampling import IIDNormalSampler

    X, Y, ls, bounds, g = _data(box=(0.0, 1.0), seed=9)
    model = SingleTaskGP(X.to(DEV), Y.to(DEV), covar_module=RBFKernel(ard_num_dims=4, lengthscale=ls)).to(DEV)
    acqf = qLogExpectedImprovement(model, best_f=Y.max().to(DEV), sampler=IIDNormalSampler(torch.Size([32]), seed=11))
    Xq = torch.rand(5, 3, 4, generator=g, dtype=torch.float64).to(DEV)
    with torch.no_grad():
        v1, v2 = acqf(Xq), acqf(Xq)
    assert torch.equal(v1, v2) and torch.isfinite(v1).all()  # base samples are drawn once and re-used
    assert acqf.sampler.base_samples.shape == (32, 1, 3)
    # a q-batch with duplicated points has a singular covariance: psd_safe_cholesky adds jitter and warns
    # (the Schur pivot of a duplicated point is rounding noise of either sign; 64 batches make a non-positive one certain)
    Xdup = torch.rand(64, 3, 4, generator=g, dtype=torch.float64).to(DEV)
    Xdup[:, 1] = Xdup[:, 0]
    Xdup[:, 2] = Xdup[:, 0]
    with warnings.catch_warnings(record=True) as ws:
        warnings.simplefilter("always")
        with torch.no_grad():
            vd = acqf(Xdup)
    assert torch.isfinite(vd).all()
    assert any(issubclass(w.category, NumericalWarning) for w in ws)


def test_error_conventions():
    from botorch_b200 import _lib
    from botorch_b200.acquisition import qLogExpectedImprovement
    from botorch_b200.models import RBFKernel, SingleTaskGP

    X, Y, ls, bounds, g = _data(box=(0.0, 1.0), seed=1)
    model = SingleTaskGP(X.to(DEV), Y.to(DEV), covar_module=RBFKernel(ard_num_dims=4, lengthscale=ls)).to(DEV)
    acqf = qLogExpectedImprovement(model, best_f=0.0)
    with pytest.raises(ValueError):
        acqf(torch.rand(4, device=DEV, dtype=torch.float64))  # fewer than 2 dims (t_batch_mode_transform)
    Xbig = torch.rand(2, 40, 4, device=DEV, dtype=torch.float64, requires_grad=True)  # q > 32: generic route, with gradient
    acqf(Xbig).sum().backward()
    assert Xbig.grad.shape == Xbig.shape and torch.isfinite(Xbig.grad).all()
    out = acqf(torch.rand(3, 2, 4, device=DEV, dtype=torch.float64))
    assert out.shape == (3,)
    assert acqf.sampler is not None and acqf.sampler.sample_shape == torch.Size([512])  # lazy default sampler


def test_float32_callers_take_the_fused_route_in_fp64():
    """fp32 optional mode: float32 models / inputs are computed in fp64 on the fused route and
    returned in the caller's dtype; against the fp64 evaluation of the same (float32-rounded) data the difference is
    float32 rounding of the output only (tolerance 1e-4 relative, as the north star states for fp32)."""
    import warnings

    from botorch_b200.acquisition import qLogNoisyExpectedImprovement
    from botorch_b200.acquisition._fused import LaunchStats
    from botorch_b200.models import RBFKernel, SingleTaskGP
    from botorch_b200.sampling import SobolQMCNormalSampler

    X, Y, ls, bounds, g = _data(box=(0.0, 1.0), seed=5)
    X32, Y32 = X.float(), Y.float()
    vals, grads = {}, {}
    for dt in (torch.float32, torch.float64):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            model = SingleTaskGP(X32.to(DEV, dt), Y32.to(DEV, dt), covar_module=RBFKernel(ard_num_dims=4, lengthscale=ls)).to(DEV)
        acqf = qLogNoisyExpectedImprovement(model, X_baseline=X32[:8].to(DEV, dt), prune_baseline=False,
                                            sampler=SobolQMCNormalSampler(torch.Size([64]), seed=1))
        Xq = torch.rand(6, 2, 4, generator=torch.Generator().manual_seed(1)).to(DEV, dt).requires_grad_(True)
        before = LaunchStats.launches
        v = acqf(Xq)
        (gr,) = torch.autograd.grad(v.sum(), Xq)
        assert LaunchStats.launches > before  # the fused kernels ran (not the generic torch-op route)
        assert v.dtype == dt and gr.dtype == dt
        vals[dt], grads[dt] = v.detach().double().cpu(), gr.double().cpu()
    assert float(((vals[torch.float32] - vals[torch.float64]).abs() / vals[torch.float64].abs()).max()) < 1e-4
    assert float((grads[torch.float32] - grads[torch.float64]).abs().max() / grads[torch.float64].abs().max()) < 1e-4


def test_qlognei_pending_points_incremental_and_joint():
    """qLogNEI `set_X_pending` (reference logei.py:393-459, 461-499): incremental=True folds the pending points into the
    baseline (== an acquisition function built on cat[X_baseline, X_pending]); incremental=False appends them along q."""
    from botorch_b200.acquisition import qLogNoisyExpectedImprovement
    from botorch_b200.models import MaternKernel, SingleTaskGP
    from botorch_b200.sampling import SobolQMCNormalSampler

    X, Y, ls, bounds, g = _data(box=(0.0, 1.0), seed=13)
    model = SingleTaskGP(X.to(DEV), Y.to(DEV), covar_module=MaternKernel(ard_num_dims=4, lengthscale=ls)).to(DEV)
    model.likelihood.noise = 2e-3
    Xb = X[:7].to(DEV)
    P = torch.rand(2, 4, generator=g, dtype=torch.float64).to(DEV)
    Xq = torch.rand(5, 2, 4, generator=g, dtype=torch.float64).to(DEV)
    mk = lambda: SobolQMCNormalSampler(torch.Size([64]), seed=3)
    inc = qLogNoisyExpectedImprovement(model, X_baseline=Xb, prune_baseline=False, sampler=mk())
    inc.set_X_pending(P)
    assert inc.X_pending is None and inc.X_baseline.shape[0] == 9
    ref = qLogNoisyExpectedImprovement(model, X_baseline=torch.cat([Xb, P]), prune_baseline=False, sampler=mk())
    with torch.no_grad():
        assert torch.equal(inc(Xq), ref(Xq))
        inc.set_X_pending(None)      # back to the plain baseline
        plain = qLogNoisyExpectedImprovement(model, X_baseline=Xb, prune_baseline=False, sampler=mk())
        assert inc.X_baseline.shape[0] == 7 and torch.equal(inc(Xq), plain(Xq))
        joint = qLogNoisyExpectedImprovement(model, X_baseline=Xb, prune_baseline=False, sampler=mk(), incremental=False)
        joint.set_X_pending(P)
        assert joint.X_pending is not None and joint.X_baseline.shape[0] == 7
        full = qLogNoisyExpectedImprovement(model, X_baseline=Xb, prune_baseline=False, sampler=mk(), incremental=False)
        assert torch.equal(joint(Xq), full(torch.cat([Xq, P.expand(5, 2, 4)], dim=-2)))
