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


def test_int8_slice_ladder_picks_more_slices_only_where_needed():
    """The probe is per fitted model: the benchmark models need the 7th forward slice at their training points; a smooth,
    well-conditioned toy model keeps 6."""
    from botorch_b200 import settings
    from botorch_b200.benchmarks import configs
    from botorch_b200.models import RBFKernel, SingleTaskGP

    with settings.contraction("int8"):
        data = configs.make_problem(configs.C2)
        strat = configs.build_model(data, DEV).prediction_strategy()
        assert (strat.contraction, strat.g_fwd) == ("int8", 7) and strat.int8_probe_error <= 2.5e-10
        g = torch.Generator().manual_seed(0)
        X = torch.rand(40, 2, generator=g, dtype=torch.float64)
        Y = torch.sin(3 * X.sum(-1, keepdim=True))
        toy = SingleTaskGP(X.to(DEV), Y.to(DEV), covar_module=RBFKernel(ard_num_dims=2, lengthscale=torch.tensor([0.8, 0.8]))).to(DEV)
        toy.likelihood.noise = 0.05
        st = toy.prediction_strategy()
        assert st.contraction == "int8" and st.g_fwd == 6
