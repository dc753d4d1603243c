"""N1 / config 4: MaxPosteriorSampling (Thompson sampling over a candidate set).  CPU: selection semantics against the
reference's documented behaviour.  GPU: joint posterior (cov + DMMA contraction + DMMA SYRK) against the oracle, the
DMMA `L z` against torch, and an end-to-end TuRBO-shaped call."""
import pytest
import torch


def test_flip_sub_unique_reference_examples():
    from botorch_b200.generation.sampling import _flip_sub_unique

    x = torch.tensor([1, 6, 4, 3, 6, 3])
    assert _flip_sub_unique(x, 3).tolist() == [3, 6, 4]  # examples of botorch/generation/utils.py:32-36
    assert _flip_sub_unique(x, 4).tolist() == [3, 6, 4, 1]
    assert _flip_sub_unique(x, 10).tolist() == [3, 6, 4, 1]


def test_maximize_samples_with_and_without_replacement():
    from botorch_b200.generation import MaxPosteriorSampling

    mps = MaxPosteriorSampling(model=None, replacement=True)
    X = torch.arange(12, dtype=torch.float64).reshape(6, 2)
    samples = torch.tensor([[0.1, 0.9, 0.3, 0.2, 0.0, 0.5], [0.2, 0.8, 0.1, 0.9, 0.0, 0.3], [0.0, 0.7, 0.6, 0.1, 0.2, 0.3]],
                           dtype=torch.float64).unsqueeze(-1)
    out = mps.maximize_samples(X, samples, num_samples=3)
    assert torch.equal(out, X[[1, 3, 1]])
    mps_nr = MaxPosteriorSampling(model=None, replacement=False)
    out_nr = mps_nr.maximize_samples(X, samples, num_samples=3)
    assert out_nr.shape == (3, 2) and len({tuple(r.tolist()) for r in out_nr}) == 3  # de-duplicated


@pytest.mark.gpu
def test_joint_posterior_and_trmm_match_oracle():
    from dataclasses import replace

    from botorch_b200.benchmarks import configs
    from oracle.harness import build_oracle

    dev = torch.device("cuda:0")
    spec = replace(configs.C3, n=512)
    data = configs.make_problem(spec)
    model = configs.build_model(data, dev)
    strat = model.prediction_strategy()
    torch.manual_seed(0)
    N = 300
    Xc = torch.rand(N, spec.d, dtype=torch.float64)
    mean, covar = strat.joint_posterior(Xc.to(dev))
    m_o, c_o = build_oracle(data).gp.posterior_mvn(Xc)
    assert float((mean.cpu() - m_o).abs().max() / m_o.abs().max()) < 1e-9
    assert float((covar.cpu() - c_o).abs().max() / c_o.abs().max()) < 1e-9
    # the public posterior API routes q > 32 through the same path
    post = model.posterior(Xc.to(dev))
    assert post.mean.shape == (N, 1) and torch.equal(post.distribution.covariance_matrix, covar)
    chol = torch.linalg.cholesky(covar)
    Z = torch.randn(64, N, device=dev, dtype=torch.float64)
    Y = strat.lower_times_samples(chol, Z)
    assert float((Y - chol @ Z.t()).abs().max() / Y.abs().max()) < 1e-12
    # a handful of samples (S <= 8) takes the memory-bound row-sweep kernel; odd N needs no padding there
    for S, Nn in ((4, N), (1, N - 1), (8, 7)):
        Zs = torch.randn(S, Nn, device=dev, dtype=torch.float64)
        Ls = chol[:Nn, :Nn].contiguous()
        Ys = strat.lower_times_samples(Ls, Zs)
        assert Ys.shape == (Nn, S)
        assert float((Ys - Ls @ Zs.t()).abs().max() / Ys.abs().max()) < 1e-12


@pytest.mark.gpu
def test_max_posterior_sampling_config4_shape():
    """TuRBO-shaped call (tutorials/turbo_1: n_candidates = min(5000, max(2000, 200 d)), replacement=False), shrunk in n."""
    from dataclasses import replace

    from botorch_b200.benchmarks import configs
    from botorch_b200.generation import MaxPosteriorSampling

    dev = torch.device("cuda:0")
    data = configs.make_problem(replace(configs.C3, n=1024))
    model = configs.build_model(data, dev)
    torch.manual_seed(0)
    Xc = torch.rand(2000, 20, dtype=torch.float64, device=dev)
    ts = MaxPosteriorSampling(model, replacement=False)
    picked = ts(Xc, num_samples=4)
    assert picked.shape == (4, 20)
    rows = {tuple(r.tolist()) for r in picked.cpu()}
    assert len(rows) == 4 and all(any(torch.equal(r, x) for x in Xc) for r in picked)
    # statistical sanity: Thompson picks concentrate where the posterior mean is high
    mean, _ = model.prediction_strategy().joint_posterior(Xc)
    ts_r = MaxPosteriorSampling(model, replacement=True)
    picked_many = ts_r(Xc, num_samples=256)
    idx = torch.tensor([int((Xc == p).all(-1).nonzero()[0]) for p in picked_many])
    assert float(mean[idx].mean()) > float(mean.mean())
    # batched X (two trust regions)
    pb = ts_r(Xc.reshape(2, 1000, 20), num_samples=3)
    assert pb.shape == (2, 3, 20)


@pytest.mark.gpu
def test_joint_posterior_at_config4_size_matches_oracle_on_row_slices():
    """BASELINE config 4 at full size: N = 5000 candidates, n = 2048.  The oracle evaluates 96-row slices of the candidate set
    (a sub-block of a joint Gaussian's covariance is the joint covariance of the sub-set), incl. the rows with the smallest
    posterior variance."""
    from dataclasses import replace

    from botorch_b200.benchmarks import configs
    from oracle.harness import build_oracle

    dev = torch.device("cuda:0")
    data = configs.make_problem(replace(configs.C3, n=2048))
    model = configs.build_model(data, dev)
    strat = model.prediction_strategy()
    g = torch.Generator().manual_seed(3)
    N = 5000
    center = data.train_X[data.train_Y.argmax()]
    Xc = (center + 0.4 * (torch.rand(N, 20, generator=g, dtype=torch.float64) - 0.5)).clamp(0.0, 1.0)
    mean, covar = strat.joint_posterior(Xc.to(dev))
    assert mean.shape == (N,) and covar.shape == (N, N)
    assert torch.equal(covar, covar.mT)      # exactly symmetric on both Gram routes (int8 slices / mirrored DMMA tiles)
    from botorch_b200 import settings

    with settings.int8_gram(False):          # DMMA SYRK-sub route for the Gram: same covariance to 1e-12 of its scale
        _, covar_dmma = strat.joint_posterior(Xc.to(dev))
    assert float((covar - covar_dmma).abs().max() / covar_dmma.abs().max()) < 1e-12
    gp = build_oracle(data).gp
    var = covar.diagonal().cpu()
    slices = [torch.arange(0, 96), torch.arange(N - 96, N), var.argsort()[:96], torch.randperm(N, generator=g)[:96]]
    for rows in slices:
        m_o, c_o = gp.posterior_mvn(Xc[rows])
        blk = covar[rows.to(dev)][:, rows.to(dev)].cpu()
        assert float((mean[rows.to(dev)].cpu() - m_o).abs().max() / m_o.abs().max()) < 1e-9
        assert float((blk - c_o).abs().max() / c_o.abs().max()) < 1e-9
        # and each variance to 1e-9 of itself
        assert float(((blk.diagonal() - c_o.diagonal()).abs() / c_o.diagonal()).max()) < 1e-9


@pytest.mark.gpu
def test_odd_candidate_count_stays_on_the_cuda_route():
    """N odd: same kernels (one padding point), same distribution as the even case on the shared candidates."""
    from dataclasses import replace

    from botorch_b200.benchmarks import configs
    from botorch_b200.generation import MaxPosteriorSampling

    dev = torch.device("cuda:0")
    data = configs.make_problem(replace(configs.C3, n=256))
    model = configs.build_model(data, dev)
    strat = model.prediction_strategy()
    torch.manual_seed(0)
    Xc = torch.rand(301, 20, dtype=torch.float64, device=dev)
    calls = []
    orig = strat.lower_times_samples
    strat.lower_times_samples = lambda chol, Z: (calls.append(chol.shape[-1]), orig(chol, Z))[1]
    picked = MaxPosteriorSampling(model, replacement=True)(Xc, num_samples=8)
    assert calls == [302] and picked.shape == (8, 20)     # the DMMA `L z` ran on the padded factor
    assert all(any(torch.equal(r, x) for x in Xc) for r in picked)
    # the N x N block that was factorised is the joint covariance of the 301 candidates
    m_pad, c_pad = strat.joint_posterior(torch.cat([Xc, Xc.mean(dim=0, keepdim=True)]))
    m_even, c_even = strat.joint_posterior(Xc[:300])
    assert float((c_pad[:300, :300] - c_even).abs().max() / c_even.abs().max()) < 1e-12
    assert float((m_pad[:300] - m_even).abs().max()) < 1e-12
