"""CPU: the restatement of L-BFGS-B that the device kernel follows (oracle/lbfgsb.py) against scipy itself -- the routine the
reference drives (botorch/optim/batched_lbfgs_b.py:365-634 steps `scipy.optimize._lbfgsb.setulb`).  Same iteration and
evaluation counts, same termination messages, iterates equal to rounding."""
import numpy as np
import pytest
from scipy.optimize import minimize

from oracle.lbfgsb import CONVERGED, STOPPED, minimize_batched


def _quad(seed, N, D):
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(N, D, D))
    Q = np.einsum("nij,nkj->nik", A, A) + 0.5 * np.eye(D)
    c = rng.normal(size=(N, D))

    def fun(X):
        d = X - c
        return 0.5 * np.einsum("ni,nij,nj->n", d, Q, d) + np.cos(3 * X).sum(-1), np.einsum("nij,nj->ni", Q, d) - 3 * np.sin(3 * X)

    return fun, rng.normal(size=(N, D))


def _rosen(X):
    f = (100 * (X[:, 1:] - X[:, :-1] ** 2) ** 2 + (1 - X[:, :-1]) ** 2).sum(-1)
    g = np.zeros_like(X)
    g[:, :-1] += -400 * X[:, :-1] * (X[:, 1:] - X[:, :-1] ** 2) - 2 * (1 - X[:, :-1])
    g[:, 1:] += 200 * (X[:, 1:] - X[:, :-1] ** 2)
    return f, g


CASES = {
    "quad6": (*_quad(0, 8, 6), -1.0, 1.5),
    "quad40": (*_quad(1, 5, 40), -0.5, 0.8),
    "rosen10": (_rosen, np.random.default_rng(3).uniform(-1.5, 1.5, size=(5, 10)), -2.0, 2.0),
    "rosen10_tight": (_rosen, np.random.default_rng(4).uniform(-0.5, 0.5, size=(5, 10)), -0.7, 0.9),
    "rosen12_unbounded": (_rosen, np.random.default_rng(5).uniform(-1.0, 1.0, size=(4, 12)), -np.inf, np.inf),
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("maxiter", [200, 7])
def test_model_tracks_scipy(name, maxiter):
    fun, x0, lo, hi = CASES[name]
    N, D = x0.shape
    xs, fs, states, rounds = minimize_batched(fun, x0, np.full(D, lo), np.full(D, hi), maxiter=maxiter)
    assert rounds == max(s.nfev for s in states)   # one batched evaluation per round, finished problems idle
    for i in range(N):
        def fi(x, i=i):
            X = x0.copy()
            X[i] = x
            f, g = fun(X)
            return float(f[i]), g[i]

        bounds = None if not np.isfinite(lo) else [(lo, hi)] * D
        ref = minimize(fi, x0[i], jac=True, method="L-BFGS-B", bounds=bounds, options={"maxiter": maxiter})
        s = states[i]
        assert (s.iter, s.nfev) == (ref.nit, ref.nfev)
        assert s.message == ref.message
        assert (s.task == CONVERGED) == bool(ref.success) and (s.task in (CONVERGED, STOPPED))
        # identical arithmetic up to the order of a few dot products: long runs may drift by the conditioning of the problem
        assert np.abs(s.x - ref.x).max() <= 1e-5 * max(1.0, np.abs(ref.x).max())
        assert abs(s.f - ref.fun) <= 1e-9 * max(1.0, abs(ref.fun))
        if maxiter == 7:
            assert np.abs(s.x - ref.x).max() <= 1e-12
