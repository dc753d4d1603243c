"""CPU tests: the oracle restatement against (a) golden vectors generated from the reference's own modules
(tests/golden/make_golden.py) and (b) the live reference modules when /root/reference is present."""
import os
import sys
import types

import pytest
import torch

from oracle import safe_math as osm
from oracle import sampling as osamp

GOLD = os.path.join(os.path.dirname(__file__), "golden")
REF = "/root/reference"


def _load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def test_reduction_chain_matches_reference_golden():
    g = _load("safe_math_chain.pt")
    z = g["z"].clone().requires_grad_(True)
    li = osm.log_fatplus(z, tau=g["tau_relu"])
    fm = osm.fatmax(li, dim=-1, tau=g["tau_max"])
    acq = osm.logmeanexp(fm, dim=0)
    (grad,) = torch.autograd.grad(acq.sum(), z)
    assert torch.equal(li.detach(), g["li"])
    assert torch.equal(fm.detach(), g["fatmax"])
    assert torch.equal(acq.detach(), g["acq"])
    assert torch.allclose(grad, g["grad"], rtol=1e-13, atol=0)
    z2 = g["z"].clone().requires_grad_(True)
    li2 = osm.log_softplus(z2, tau=g["tau_relu"])
    fm2 = osm.smooth_amax(li2, dim=-1, tau=g["tau_max"])
    acq2 = osm.logmeanexp(fm2, dim=0)
    assert torch.equal(li2.detach(), g["li_nofat"])
    assert torch.equal(fm2.detach(), g["smooth_amax"])
    assert torch.equal(acq2.detach(), g["acq_nofat"])


def test_inf_handling_matches_reference_golden():
    g = _load("safe_math_inf.pt")
    x = g["x"]
    assert torch.equal(osm.fatmax(x, dim=-1, tau=1e-2), g["fatmax"])
    assert torch.equal(osm.logsumexp(x, dim=-1), g["logsumexp"])
    assert torch.equal(osm.logmeanexp(x, dim=-1), g["logmeanexp"])
    # empty reductions are -inf (test/utils/test_safe_math.py)
    assert osm.logsumexp(torch.zeros(3, 0, dtype=torch.float64), dim=-1).eq(-torch.inf).all()


def test_qmc_draws_match_reference_golden():
    g = _load("qmc_draws.pt")
    for (d, n, seed) in [(4, 16, 1234), (24, 32, 1234), (8, 8, 7)]:
        assert torch.equal(osamp.draw_sobol_normal_samples(d, n, torch.float64, seed), g[f"normal_d{d}_n{n}_s{seed}"])
    bounds = torch.stack([torch.zeros(6, dtype=torch.float64), torch.ones(6, dtype=torch.float64)])
    assert torch.equal(osamp.draw_sobol_samples(bounds, 8, 4, seed=0), g["sobol_n8_q4_d6_s0"])


def test_product_host_draws_match_reference_golden():
    """The product's host-side Sobol helpers (botorch_b200/utils/sampling.py) are bit-identical to the reference."""
    from botorch_b200.utils import sampling as ps

    g = _load("qmc_draws.pt")
    for (d, n, seed) in [(4, 16, 1234), (24, 32, 1234), (8, 8, 7)]:
        assert torch.equal(ps.draw_sobol_normal_samples(d=d, n=n, dtype=torch.float64, seed=seed), g[f"normal_d{d}_n{n}_s{seed}"])
    bounds = torch.stack([torch.zeros(6, dtype=torch.float64), torch.ones(6, dtype=torch.float64)])
    assert torch.equal(ps.draw_sobol_samples(bounds=bounds, n=8, q=4, seed=0), g["sobol_n8_q4_d6_s0"])
    b2 = torch.tensor([[-1.0, 0.0, 2.0], [1.0, 5.0, 2.5]], dtype=torch.float64)
    assert torch.allclose(ps.draw_sobol_samples(bounds=b2, n=4, q=2, seed=11), g["sobol_n4_q2_d3_s11"], rtol=1e-15, atol=0)
    w = torch.tensor([0.3, -1.2, 0.8, 2.5, 2.4, -0.1], dtype=torch.float64)
    with ps.manual_seed(5):
        assert torch.equal(ps.boltzmann_sample(w, num_samples=3, eta=2.0), g["boltzmann_idx"])


def test_product_safe_math_matches_reference_golden():
    from botorch_b200.utils import safe_math as psm

    g = _load("safe_math_chain.pt")
    z = g["z"]
    fm = psm.fatmax(psm.log_fatplus(z, tau=g["tau_relu"]), dim=-1, tau=g["tau_max"])
    assert torch.equal(psm.logmeanexp(fm, dim=0), g["acq"])
    gi = _load("safe_math_inf.pt")
    assert torch.equal(psm.fatmax(gi["x"], dim=-1, tau=1e-2), gi["fatmax"])
    assert torch.equal(psm.logsumexp(gi["x"], dim=-1), gi["logsumexp"])


def test_oracle_gp_path_regression_vectors():
    """Oracle-minted vectors (NOT reference outputs; the GP boundary is 'parity unpinned')."""
    from oracle.acquisition import OracleQLogEI, OracleQLogNEI, value_and_grad
    from oracle.gp import OracleGP

    g = _load("oracle_path_small.pt")
    for kern in ("rbf", "matern52"):
        gp = OracleGP(g["X"], g["Y"], g["ls"], torch.tensor(1e-3, dtype=torch.float64), kernel=kern,
                      outputscale=1.7 if kern == "matern52" else None)
        mean, cov = gp.posterior_mvn(g["Xq"])
        assert torch.allclose(mean, g[kern]["mean"], rtol=1e-10, atol=1e-12)
        assert torch.allclose(cov, g[kern]["cov"], rtol=1e-9, atol=1e-12)
        v1, g1 = value_and_grad(OracleQLogEI(gp, g["Y"].max(), 64, 1234), g["Xq"])
        v2, g2 = value_and_grad(OracleQLogNEI(gp, g["X"][:5], 64, 1234), g["Xq"])
        assert torch.allclose(v1, g[kern]["qlogei"], rtol=1e-9)
        assert torch.allclose(v2, g[kern]["qlognei"], rtol=1e-9)
        assert torch.allclose(g1, g[kern]["qlogei_grad"], rtol=1e-7, atol=1e-9)
        assert torch.allclose(g2, g[kern]["qlognei_grad"], rtol=1e-7, atol=1e-9)


def test_oracle_posterior_identities():
    """Independent anchors for the restated GP arithmetic: the explicit `K - K_* (K+s2 I)^{-1} K_*^T` formula, the
    interpolation limit, and (q=1) analytic log-EI vs the MC value."""
    from oracle.gp import OracleGP, rbf_forward

    torch.manual_seed(3)
    n, d = 30, 2
    X = torch.rand(n, d, dtype=torch.float64)
    Y = torch.cos(4 * X[:, :1]) + X[:, 1:]
    ls = torch.tensor([0.3, 0.5], dtype=torch.float64)
    gp = OracleGP(X, Y, ls, torch.tensor(1e-4, dtype=torch.float64), standardize=False)
    Xq = torch.rand(4, 3, d, dtype=torch.float64)
    mean, cov = gp.posterior_mvn(Xq)
    K = rbf_forward(X, X, ls) + 1e-4 * torch.eye(n, dtype=torch.float64)
    for b in range(4):
        Ks = rbf_forward(Xq[b], X, ls)
        ref_mean = Ks @ torch.linalg.solve(K, Y.squeeze(-1))
        ref_cov = rbf_forward(Xq[b], Xq[b], ls) - Ks @ torch.linalg.solve(K, Ks.T)
        assert torch.allclose(mean[b], ref_mean, atol=1e-9)
        assert torch.allclose(cov[b], ref_cov, atol=1e-9)
    m_tr, c_tr = gp.posterior_mvn(X[:5].unsqueeze(0))
    assert torch.allclose(m_tr[0], Y[:5, 0], atol=1e-2)
    assert (c_tr[0].diagonal() < 2e-4).all()


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present on this machine")
def test_oracle_safe_math_equals_live_reference():
    for name, sub in (("botorch", ""), ("botorch.utils", "utils"), ("botorch.sampling", "sampling")):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            mod.__path__ = [os.path.join(REF, "botorch", sub)]
            sys.modules[name] = mod
    import importlib

    rsm = importlib.import_module("botorch.utils.safe_math")
    rus = importlib.import_module("botorch.utils.sampling")
    torch.manual_seed(11)
    for scale in (1e-8, 1e-3, 1.0, 50.0):
        z = torch.randn(33, 7, 5, dtype=torch.float64) * scale
        for tau in (1e-6, 1e-2, 1.0):
            assert torch.equal(osm.log_fatplus(z, tau=tau), rsm.log_fatplus(z, tau=tau))
            assert torch.equal(osm.log_softplus(z, tau=tau), rsm.log_softplus(z, tau=tau))
        li = rsm.log_fatplus(z, tau=1e-6)
        assert torch.equal(osm.fatmax(li, dim=-1, tau=1e-2), rsm.fatmax(li, dim=-1, tau=1e-2))
        assert torch.equal(osm.smooth_amax(li, dim=-1, tau=1e-2), rsm.smooth_amax(li, dim=-1, tau=1e-2))
        assert torch.equal(osm.logmeanexp(li, dim=0), rsm.logmeanexp(li, dim=0))
    assert torch.equal(osamp.draw_sobol_normal_samples(11, 64, torch.float64, 99),
                       rus.draw_sobol_normal_samples(d=11, n=64, dtype=torch.float64, seed=99))
    # the (r+q)-dim draw's first r columns differ from the r-dim draw: the overwrite is load-bearing
    full = osamp.draw_sobol_normal_samples(24, 64, torch.float64, 1234)
    base = osamp.draw_sobol_normal_samples(16, 64, torch.float64, 1234)
    assert not torch.allclose(full[:, :16], base)
    assert torch.equal(osamp.qlognei_base_samples(64, 16, 8, 1234).view(64, 24)[:, :16], base)
