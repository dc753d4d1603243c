"""GPU: parity AT and NEAR the training points, at the full benchmark sizes, in both contraction modes (VERDICT r01, next
round item 1a).  There the posterior variance has collapsed by four orders of magnitude (K_** - A A^T is a difference of
nearly equal numbers), which is where a fixed-point contraction loses digits first: the int8 mode must have picked its
slice count so that the north-star bar still holds (posterior variance and acquisition values within 1e-9 of the oracle).
q-batches are made of training points that are NOT baseline points: a q-batch point identical to a baseline point makes the
conditional covariance exactly singular, and then the reference's own result is decided by rounding (jitter or not)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
DELTAS = (0.0, 1e-6, 1e-3)


def _near_train(data, count, delta, seed=0):
    g = torch.Generator().manual_seed(seed)
    n = data.train_X.shape[0]
    base_idx = set()
    if data.X_baseline is not None:
        d2 = torch.cdist(data.X_baseline, data.train_X)
        base_idx = set(d2.argmin(dim=1).tolist())
    idx = [i for i in torch.linspace(0, n - 1, 3 * count).round().long().tolist() if i not in base_idx][:count]
    T = data.train_X[idx]
    dirs = torch.randn(T.shape, generator=g, dtype=torch.float64)
    dirs = dirs / dirs.norm(dim=-1, keepdim=True)
    return (T + delta * dirs).clamp(0.0, 1.0)


@pytest.mark.parametrize("mode", ["int8", "dmma"])
@pytest.mark.parametrize("cfg", ["C2", "C3"])
def test_posterior_and_acquisition_at_training_points_full_size(cfg, mode):
    from botorch_b200 import settings
    from botorch_b200.benchmarks import configs
    from oracle.acquisition import value_and_grad
    from oracle.harness import build_oracle

    spec = configs.CONFIGS[cfg]
    data = configs.make_problem(spec)
    orc = build_oracle(data)
    with settings.contraction(mode):
        model = configs.build_model(data, DEV)
        strat = model.prediction_strategy()
        acqf = configs.build_acqf(data, model)
        assert strat.contraction == mode    # no silent fall-back on the benchmark problems
        if mode == "int8":
            assert strat.int8_probe_error <= strat.INT8_PROBE_TOL and strat.g_fwd in strat.G_FWD_LADDER
            assert strat.int8_var_ratio_limit is not None and 0.0 < strat.int8_var_ratio_limit < 0.05
        from botorch_b200.acquisition._fused import RerouteStats

        RerouteStats.q_batches = RerouteStats.calls = 0
        for delta in DELTAS:
            P = _near_train(data, 48, delta)
            post = model.posterior(P.unsqueeze(1).to(DEV))
            m_o, c_o = orc.gp.posterior_mvn(P.unsqueeze(1))
            m, v = post.mean.reshape(-1).cpu(), post.variance.reshape(-1).cpu()
            assert float(((m - m_o.reshape(-1)).abs() / m_o.reshape(-1).abs().clamp_min(1e-300)).max()) < 1e-9
            relv = ((v - c_o.reshape(-1)).abs() / c_o.reshape(-1)).max()
            assert float(relv) < 1e-9, (cfg, mode, delta, float(relv), float(c_o.min()))
            # acquisition values of q-batches of (displaced) training points
            q = spec.q
            Xq = P[: (P.shape[0] // q) * q].reshape(-1, q, spec.d)[:5]
            v_o, g_o = value_and_grad(orc, Xq)
            Xg = Xq.to(DEV).requires_grad_(True)
            val = acqf(Xg)
            (gr,) = torch.autograd.grad(val.sum(), Xg)
            rel = ((val.detach().cpu() - v_o).abs() / v_o.abs()).max()
            assert float(rel) < 1e-9, (cfg, mode, delta, float(rel))
            assert float((gr.cpu() - g_o).abs().max() / g_o.abs().max()) < 1e-6, (cfg, mode, delta)
        # the int8 mode holds the bar at the training points BECAUSE it recognises them (variance-collapse byte of the status
        # word) and evaluates those q-batches through the FP64 contraction; ordinary Sobol points stay on the int8 path
        if mode == "int8":
            assert RerouteStats.q_batches > 0
            RerouteStats.q_batches = RerouteStats.calls = 0
            Xs = configs.eval_points(data, 256).to(DEV)
            with torch.no_grad():
                acqf(Xs)
            assert RerouteStats.q_batches <= 2     # (the 99.9 % quantile of the sweep's collapse byte sits at the limit)
        else:
            assert RerouteStats.q_batches == 0


def test_int8_slice_ladder_and_collapse_limit_per_model():
    """The probe is per fitted model.  The benchmark models keep 6 forward slices for ordinary points (their variance-collapse
    limit is ~1 % of the prior; the training points lie below it and take the FP64 route)."""
    from botorch_b200 import settings
    from botorch_b200.benchmarks import configs
    from botorch_b200.models import RBFKernel, SingleTaskGP

    with settings.contraction("int8"):
        data = configs.make_problem(configs.C2)
        strat = configs.build_model(data, DEV).prediction_strategy()
        assert strat.contraction == "int8" and strat.g_fwd == 6 and strat.int8_probe_error <= 2.5e-10
        assert 1e-4 < strat.int8_var_ratio_limit < 0.05 and 17 <= strat.int8_var_byte_limit <= 53
        # the posterior API applies the same rule: variances AT training points to 1e-9 although G = 6 alone gives ~4e-9
        from oracle.harness import build_oracle

        P = data.train_X[:64].unsqueeze(1)
        v = configs.build_model(data, DEV).posterior(P.to(DEV)).variance.reshape(-1).cpu()
        _, c_o = build_oracle(data).gp.posterior_mvn(P)
        assert float(((v - c_o.reshape(-1)).abs() / c_o.reshape(-1)).max()) < 1e-9
        # the CUDA-graph rounds of the device optimiser cannot re-route: they run the most accurate slice counts
        mv = strat.max_slices_view()
        assert (mv.g_fwd, mv.g_bwd) == (7, 7) and mv.int8_var_byte_limit is None and mv.desc.g_fwd == 7
        g = torch.Generator().manual_seed(0)
        X = torch.rand(40, 2, generator=g, dtype=torch.float64)
        Y = torch.sin(3 * X.sum(-1, keepdim=True))
        toy = SingleTaskGP(X.to(DEV), Y.to(DEV), covar_module=RBFKernel(ard_num_dims=2, lengthscale=torch.tensor([0.8, 0.8]))).to(DEV)
        toy.likelihood.noise = 0.05
        st = toy.prediction_strategy()
        assert st.contraction == "int8" and st.g_fwd == 6
