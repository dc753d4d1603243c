"""CPU tests of the host-side optimisation callers with a mock acquisition function (reference strategy:
test/optim/test_initializers.py, test/generation/test_gen.py, test/optim/test_batched_lbfgs_b.py)."""
import numpy as np
import pytest
import torch
from scipy.optimize import minimize


class QuadraticAcqf(torch.nn.Module):
    """acq(X) = -sum_q ||x - c||^2 : maximum at x = c (differentiable, batch independent)."""

    X_pending = None

    def __init__(self, c):
        super().__init__()
        self.c = c

    def set_X_pending(self, X):
        self.X_pending = X

    def forward(self, X):
        return -((X - self.c) ** 2).sum(dim=(-1, -2))


def test_batched_lbfgsb_matches_scipy_per_problem():
    from botorch_b200.optim.batched_lbfgs_b import fmin_l_bfgs_b_batched

    rng = np.random.default_rng(0)
    N, D = 7, 5
    A = rng.normal(size=(N, D, D))
    Q = np.einsum("nij,nkj->nik", A, A) + 0.5 * np.eye(D)
    c = rng.normal(size=(N, D))
    calls = []

    def func(X, batch_indices):
        calls.append(len(batch_indices))
        idx = np.array(batch_indices)
        diff = X - c[idx]
        f = 0.5 * np.einsum("ni,nij,nj->n", diff, Q[idx], diff) + np.cos(X).sum(-1)
        g = np.einsum("nij,nj->ni", Q[idx], diff) - np.sin(X)
        return f, g

    x0 = rng.normal(size=(N, D))
    bounds = [(-1.0, 1.5)] * D
    xs, fs, results = fmin_l_bfgs_b_batched(func, x0, bounds=bounds, maxiter=200, pass_batch_indices=True)
    assert max(calls) == N and min(calls) >= 1  # evaluations are batched, the active set shrinks
    for i in range(N):
        def fi(x, i=i):
            f, g = func(x[None], [i])
            return float(f[0]), g[0]
        ref = minimize(fi, x0[i], jac=True, method="L-BFGS-B", bounds=bounds, options={"maxiter": 200})
        assert np.array_equal(ref.x, xs[i])  # identical iterates => bit-identical solutions
        assert ref.fun == fs[i] and ref.nit == results[i].nit


def test_batched_lbfgsb_propagates_errors():
    from botorch_b200.optim.batched_lbfgs_b import fmin_l_bfgs_b_batched

    def bad(X, batch_indices):
        raise RuntimeError("boom")

    with pytest.raises(RuntimeError, match="boom"):
        fmin_l_bfgs_b_batched(bad, np.zeros((3, 2)), pass_batch_indices=True)


def test_gen_candidates_scipy_reaches_optimum_and_clamps():
    from botorch_b200.exceptions import BotorchError
    from botorch_b200.generation import gen_candidates_scipy

    c = torch.tensor([0.3, 0.7, 0.5], dtype=torch.float64)
    acqf = QuadraticAcqf(c)
    ics = torch.rand(6, 2, 3, dtype=torch.float64)
    cand, val = gen_candidates_scipy(ics, acqf, lower_bounds=0.0, upper_bounds=1.0)
    assert cand.shape == ics.shape and val.shape == (6,)
    assert torch.allclose(cand, c.expand_as(cand), atol=1e-5)
    ics_in = ics.clone()
    ics_in[..., 1] *= 0.6
    cand2, _ = gen_candidates_scipy(ics_in, acqf, lower_bounds=0.0, upper_bounds=torch.tensor([1.0, 0.6, 1.0]))
    assert torch.allclose(cand2[..., 1], torch.full_like(cand2[..., 1], 0.6), atol=1e-6)
    with pytest.raises(BotorchError):
        gen_candidates_scipy(ics + 2.0, acqf, lower_bounds=0.0, upper_bounds=1.0)


def test_nan_gradient_raises_optimization_gradient_error():
    from botorch_b200.exceptions import OptimizationGradientError
    from botorch_b200.generation import gen_candidates_scipy

    class NanGrad(QuadraticAcqf):
        def forward(self, X):
            return super().forward(X) * torch.tensor(float("nan"), dtype=X.dtype)

    with pytest.raises(OptimizationGradientError):
        gen_candidates_scipy(torch.rand(2, 1, 3, dtype=torch.float64), NanGrad(torch.zeros(3, dtype=torch.float64)),
                             lower_bounds=0.0, upper_bounds=1.0)


def test_initial_conditions_selection_semantics():
    from botorch_b200.exceptions import BadInitialCandidatesWarning
    from botorch_b200.optim import gen_batch_initial_conditions, initialize_q_batch, initialize_q_batch_topn

    bounds = torch.stack([torch.zeros(3, dtype=torch.float64), torch.ones(3, dtype=torch.float64)])
    acqf = QuadraticAcqf(torch.tensor([0.2, 0.2, 0.9], dtype=torch.float64))
    torch.manual_seed(0)
    ics = gen_batch_initial_conditions(acqf, bounds, q=2, num_restarts=4, raw_samples=64, options={"seed": 3})
    assert ics.shape == (4, 2, 3) and (ics >= 0).all() and (ics <= 1).all()
    torch.manual_seed(0)
    ics2 = gen_batch_initial_conditions(acqf, bounds, q=2, num_restarts=4, raw_samples=64, options={"seed": 3, "init_batch_limit": 7})
    assert torch.equal(ics, ics2)  # chunking the sweep does not change the selection
    X = torch.rand(10, 1, 3, dtype=torch.float64)
    vals = torch.arange(10, dtype=torch.float64)
    torch.manual_seed(1)
    Xs, vs = initialize_q_batch(X, vals, n=3, eta=1.0)
    assert 9.0 in vs.tolist()  # the arg-max is always included (initializers.py:1028-1030)
    Xt, vt = initialize_q_batch_topn(X, vals, n=3)
    assert vt.tolist() == [9.0, 8.0, 7.0]
    with pytest.warns(BadInitialCandidatesWarning):
        initialize_q_batch(X, torch.ones(10, dtype=torch.float64), n=3)
    with pytest.raises(RuntimeError):
        initialize_q_batch(X, vals, n=11)


def test_optimize_acqf_joint_and_sequential():
    from botorch_b200.optim import optimize_acqf

    bounds = torch.stack([torch.zeros(2, dtype=torch.float64), torch.ones(2, dtype=torch.float64)])
    c = torch.tensor([0.25, 0.6], dtype=torch.float64)
    acqf = QuadraticAcqf(c)
    torch.manual_seed(0)
    cand, val = optimize_acqf(acqf, bounds, q=2, num_restarts=3, raw_samples=32, options={"seed": 0})
    assert cand.shape == (2, 2) and torch.allclose(cand, c.expand(2, 2), atol=1e-5) and val.ndim == 0
    allc, allv = optimize_acqf(acqf, bounds, q=2, num_restarts=3, raw_samples=32, options={"seed": 0}, return_best_only=False)
    assert allc.shape == (3, 2, 2) and allv.shape == (3,)
    cs, vs = optimize_acqf(acqf, bounds, q=2, num_restarts=3, raw_samples=32, options={"seed": 0}, sequential=True)
    assert cs.shape == (2, 2) and vs.shape == (2,) and acqf.X_pending is None


@pytest.mark.parametrize("driver", ["direct", "threads"])
def test_batched_lbfgsb_drivers_agree_with_scipy_on_bound_types(driver, monkeypatch):
    """Unbounded, one-sided, per-problem (N x D x 2) bounds, an iteration cap and a halting callback: both drivers must
    reproduce scipy.optimize.minimize(method='L-BFGS-B') per problem (bit-identical x, f, nit, nfev)."""
    from botorch_b200.optim import batched_lbfgs_b as B

    rng = np.random.default_rng(3)
    N, D = 5, 4
    c = rng.normal(size=(N, D))

    def func(X, batch_indices):
        idx = np.array(batch_indices)
        d = X - c[idx]
        return (d**4).sum(-1) + 0.5 * (d**2).sum(-1) + np.sin(X).sum(-1), 4 * d**3 + d + np.cos(X)

    def run(**kw):
        if driver == "threads":
            return B._run_threads(func, kw["x0"], kw.get("bounds"), kw.get("maxiter", 15000), 10, 2.2204460492503131e-09, 1e-5,
                                  20, 15000, kw.get("callback"), True)
        drv = B._direct_driver()
        assert drv is not None
        return B._run_direct(drv[0], drv[1], func, kw["x0"], kw.get("bounds"), kw.get("maxiter", 15000), 10,
                             2.2204460492503131e-09, 1e-5, 20, 15000, kw.get("callback"), True)

    x0 = rng.normal(size=(N, D))
    per_problem = np.stack([np.stack([c[i] - 0.3 - 0.1 * i, c[i] + 0.2 + 0.05 * i], axis=-1) for i in range(N)])
    cases = {"unbounded": None, "one_sided": [(None, 0.4), (-0.2, None), (None, None), (-1.0, 1.0)], "per_problem": per_problem}
    for name, bnds in cases.items():
        xs, fs, res = run(x0=x0, bounds=bnds, maxiter=60)
        for i in range(N):
            bi = None if bnds is None else (bnds[i] if name == "per_problem" else bnds)
            if isinstance(bi, np.ndarray):
                bi = [tuple(r) for r in bi]
            ref = minimize(lambda x, i=i: tuple(v[0] for v in func(x[None], [i])), x0[i], jac=True, method="L-BFGS-B",
                           bounds=bi, options={"maxiter": 60})
            assert np.array_equal(ref.x, xs[i]) and ref.fun == fs[i], (name, i)
            assert ref.nit == res[i].nit and ref.nfev == res[i].nfev and ref.status == res[i].status, (name, i)
    # iteration cap: status 1 like scipy
    xs, fs, res = run(x0=x0, bounds=None, maxiter=2)
    assert all(r.nit == 2 and r.status == 1 and not r.success for r in res)
    # a callback that halts after the third iterate of every problem
    seen = []

    def cb(xk):
        seen.append(xk.copy())
        if len(seen) % 3 == 0 and driver == "direct":
            raise StopIteration

    if driver == "direct":
        xs, fs, res = run(x0=x0[:1], bounds=None, callback=cb)
        assert res[0].nit == 3 and not res[0].success
    # infeasible bounds are rejected up front
    with pytest.raises(ValueError):
        run(x0=x0, bounds=[(1.0, 0.0)] * D)


def test_sobol_engine_cache_keeps_fresh_engines():
    """Seeded engines are cached un-advanced (the device draw only reads their state); unseeded ones are never cached; the
    closed form the device kernel evaluates reproduces torch's sequence (checked here in integer arithmetic on the host)."""
    from torch.quasirandom import SobolEngine

    from botorch_b200.utils.sampling import _ENGINES, _fresh_engine, draw_sobol_samples

    a, b = _fresh_engine(6, 42), _fresh_engine(6, 42)
    assert a is b and a.num_generated == 0 and (6, 42) in _ENGINES
    assert _fresh_engine(6, None) is not _fresh_engine(6, None)
    for i in range(40):
        _fresh_engine(3, 1000 + i)
    assert len(_ENGINES) <= 16
    # host restatement of mcacq_sobol_draw: x_k = shift XOR_{b in gray(k)} sobolstate[:, b], scaled by 2^-30; row 0 is the
    # engine's (default-dtype rounded) first point
    eng = SobolEngine(9, scramble=True, seed=7)
    k = torch.arange(0, 700, dtype=torch.int64)
    gray = k ^ (k >> 1)
    acc = eng.shift.unsqueeze(0).expand(700, -1).clone()
    for bit in range(30):
        m = ((gray >> bit) & 1).bool()
        acc[m] ^= eng.sobolstate[:, bit]
    pts = acc.to(torch.float64) * 2.0 ** -30
    pts[0] = eng._first_point.to(torch.float64).reshape(-1)
    assert torch.equal(pts, SobolEngine(9, scramble=True, seed=7).draw(700, dtype=torch.float64))
    # CPU bounds keep the reference's host path
    bounds = torch.tensor([[0.0, -1.0], [2.0, 1.0]], dtype=torch.float64)
    x = draw_sobol_samples(bounds=bounds, n=16, q=3, seed=5)
    ref = SobolEngine(6, scramble=True, seed=5).draw(16, dtype=torch.float64).view(16, 3, 2) * (bounds[1] - bounds[0]) + bounds[0]
    assert torch.equal(x, ref)
