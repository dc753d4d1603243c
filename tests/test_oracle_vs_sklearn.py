"""CPU: the GP part of the oracle (`oracle/gp.py`, a from-memory restatement of gpytorch's exact prediction path -- "parity
unpinned" against gpytorch itself, which is not installable here) against an INDEPENDENT third-party implementation of the
same published algorithm: scikit-learn's `GaussianProcessRegressor` (Rasmussen & Williams Alg. 2.1) with fixed
hyper-parameters, ARD RBF / Matern-5/2 kernels, ConstantKernel as the outputscale, `alpha` as the (homo- or
heteroskedastic) noise.  Pins posterior mean and full q x q covariance on the ORIGINAL outcome scale (Normalize and
Standardize applied by hand on the scikit-learn side, as botorch/models/transforms/input.py:541-554 and outcome.py:340-352,
479-511 define them), including exactly AT training points where the variance has collapsed."""
import numpy as np
import pytest
import torch

sk_gp = pytest.importorskip("sklearn.gaussian_process")
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern  # noqa: E402

from oracle.gp import OracleGP  # noqa: E402


def _problem(seed, n, d, kernel, scale, normalize, fixed_noise):
    g = torch.Generator().manual_seed(seed)
    lo, hi = (-2.0, 3.0) if normalize else (0.0, 1.0)
    X = lo + (hi - lo) * torch.rand(n, d, generator=g, dtype=torch.float64)
    Y = torch.sin(2.5 * (X - lo).sum(-1, keepdim=True) / (hi - lo) / d ** 0.5) + 0.1 * torch.randn(n, 1, generator=g, dtype=torch.float64)
    ls = (0.15 + 0.25 * torch.rand(d, generator=g, dtype=torch.float64)) * d ** 0.5
    noise = (2e-3 + 5e-3 * torch.rand(n, generator=g, dtype=torch.float64)) if fixed_noise else torch.tensor(4e-3, dtype=torch.float64)
    bounds = torch.tensor([[lo] * d, [hi] * d], dtype=torch.float64)
    return X, Y, ls, noise, bounds, g


@pytest.mark.parametrize("seed,n,d,q,kernel,scale,normalize,fixed_noise", [
    (0, 40, 3, 4, "rbf", None, False, False),
    (1, 120, 6, 5, "matern52", None, False, False),
    (2, 200, 20, 8, "matern52", 1.7, True, False),
    (3, 64, 2, 3, "rbf", 0.6, True, True),
    (4, 300, 10, 8, "rbf", None, False, True),
    (5, 17, 1, 2, "matern52", 2.5, False, False),
])
def test_oracle_posterior_matches_scikit_learn(seed, n, d, q, kernel, scale, normalize, fixed_noise):
    X, Y, ls, noise, bounds, g = _problem(seed, n, d, kernel, scale, normalize, fixed_noise)
    mean_const = 0.3 if seed % 2 else 0.0
    gp = OracleGP(X, Y, ls, noise, kernel=kernel, outputscale=scale, mean_constant=mean_const,
                  norm_offset=bounds[0] if normalize else None, norm_coef=(bounds[1] - bounds[0]) if normalize else None)
    lo, hi = float(bounds[0, 0]), float(bounds[1, 0])
    Xq = lo + (hi - lo) * torch.rand(7, q, d, generator=g, dtype=torch.float64)
    Xq[0, 0] = X[3]                     # exactly at a training point
    Xq[1, 1] = X[5] + 1e-6              # and next to one
    # ---- scikit-learn side: transforms by hand, zero-mean GP on the standardised residual
    tf = (lambda Z: (Z - bounds[0]) / (bounds[1] - bounds[0])) if normalize else (lambda Z: Z)
    m, s = Y.mean(), Y.std()            # torch.std is the unbiased estimator Standardize uses
    y_std = ((Y - m) / s).reshape(-1) - mean_const
    base = (RBF if kernel == "rbf" else (lambda length_scale: Matern(length_scale=length_scale, nu=2.5)))(length_scale=ls.numpy())
    k = ConstantKernel(scale, constant_value_bounds="fixed") * base if scale else base
    reg = sk_gp.GaussianProcessRegressor(kernel=k, alpha=noise.numpy() if fixed_noise else float(noise), optimizer=None,
                                         normalize_y=False).fit(tf(X).numpy(), y_std.numpy())
    mean_o, cov_o = gp.posterior_mvn(Xq)
    for i in range(Xq.shape[0]):
        mu, cov = reg.predict(tf(Xq[i]).numpy(), return_cov=True)
        mu = float(m) + float(s) * (mu + mean_const)
        cov = float(s) ** 2 * cov
        assert np.abs(mean_o[i].numpy() - mu).max() <= 1e-9 * max(1.0, np.abs(mu).max())
        # both sides form K** - V^T V in fp64: agreement is limited by that cancellation (eps * prior variance), so the
        # bar is relative to the prior variance, and 1e-9 of each entry wherever the variance has not collapsed
        prior = float(s) ** 2 * (scale or 1.0)
        assert np.abs(cov_o[i].numpy() - cov).max() <= 1e-11 * prior
        big = np.abs(cov) > 1e-3 * prior
        assert (np.abs(cov_o[i].numpy() - cov)[big] <= 1e-9 * np.abs(cov)[big]).all()


def test_oracle_caches_match_scikit_learn_factor():
    """`mean_cache` = (K + noise)^-1 (y - mu) and the train Cholesky factor against scikit-learn's `alpha_` and `L_`."""
    X, Y, ls, noise, bounds, g = _problem(7, 150, 5, "matern52", None, False, False)
    gp = OracleGP(X, Y, ls, noise, kernel="matern52")
    _, L, mean_cache, covar_cache, m, s = gp.caches()
    reg = sk_gp.GaussianProcessRegressor(kernel=Matern(length_scale=ls.numpy(), nu=2.5), alpha=float(noise), optimizer=None).fit(
        X.numpy(), ((Y - Y.mean()) / Y.std()).reshape(-1).numpy())
    assert np.abs(L.numpy() - reg.L_).max() <= 1e-10
    assert np.abs(mean_cache.numpy() - reg.alpha_).max() <= 1e-8 * np.abs(reg.alpha_).max()
    # covar_cache = L^-T: covar_cache^T L = I
    assert np.abs(covar_cache.numpy().T @ reg.L_ - np.eye(150)).max() <= 1e-10
