"""CPU tests of the optimiser host semantics added in round 2 (reference strategy: test/optim/test_initializers.py,
test/optim/test_optimize.py, test/optim/test_batched_lbfgs_b.py, test/optim/utils/test_timeout.py):
non-negative initial-condition heuristic, `sample_around_best`, retry on `OptimizationWarning`, timeout,
post-processing before the arg-max, and the reference's OWN `optim/batched_lbfgs_b.py` as the checker of our drivers
(imported by file path when /root/reference exists, i.e. in the build container)."""
import importlib.util
import os
import sys
import types
import warnings

import numpy as np
import pytest
import torch

REF = "/root/reference/botorch/optim/batched_lbfgs_b.py"


class QuadraticAcqf(torch.nn.Module):
    X_pending = None

    def __init__(self, c):
        super().__init__()
        self.c = c

    def set_X_pending(self, X):
        self.X_pending = X

    def forward(self, X):
        return -((X - self.c) ** 2).sum(dim=(-1, -2))


def _reference_batched():
    """The reference's batched L-BFGS-B module, loaded by path with a one-function shim for its relative import."""
    name = "_ref_optim_pkg"
    pkg = types.ModuleType(name)
    pkg.__path__ = []
    utils = types.ModuleType(name + ".utils")

    def check_scipy_version_at_least(minor: int, major: int = 1) -> bool:
        import scipy

        ma, mi = (int(v) for v in scipy.__version__.split(".")[:2])
        return (ma, mi) >= (major, minor)

    utils.check_scipy_version_at_least = check_scipy_version_at_least
    sys.modules[name], sys.modules[name + ".utils"] = pkg, utils
    spec = importlib.util.spec_from_file_location(name + ".batched_lbfgs_b", REF)
    mod = importlib.util.module_from_spec(spec)
    mod.__package__ = name
    spec.loader.exec_module(mod)
    return mod


def _problem(seed=0, N=9, D=6):
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(N, D, D))
    Q = np.einsum("nij,nkj->nik", A, A) + 0.5 * np.eye(D)
    c = rng.normal(size=(N, D))

    def func(X, batch_indices):
        idx = np.array(batch_indices)
        diff = X - c[idx]
        f = 0.5 * np.einsum("ni,nij,nj->n", diff, Q[idx], diff) + np.cos(3 * X).sum(-1)
        g = np.einsum("nij,nj->ni", Q[idx], diff) - 3 * np.sin(3 * X)
        return f, g

    return func, rng.normal(size=(N, D)), [(-1.0, 1.5)] * D


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("driver", ["direct", "threads"])
@pytest.mark.parametrize("maxiter", [200, 7])
def test_drivers_match_the_reference_batched_lbfgsb(driver, maxiter, monkeypatch):
    """Identical iterates and termination against the reference's scipy-only module, for both of our drivers."""
    import botorch_b200.optim.batched_lbfgs_b as ours

    monkeypatch.setenv("BOTORCH_B200_LBFGSB", driver)
    monkeypatch.setattr(ours, "_DIRECT", None)
    ref = _reference_batched()
    func, x0, bounds = _problem()
    trace_ref, trace_ours = [], []
    xr, fr, rr = ref.fmin_l_bfgs_b_batched(lambda X, batch_indices: (trace_ref.append(X.copy()), func(X, batch_indices))[1],
                                           x0, bounds=bounds, maxiter=maxiter, pass_batch_indices=True)
    xo, fo, ro = ours.fmin_l_bfgs_b_batched(lambda X, batch_indices: (trace_ours.append(X.copy()), func(X, batch_indices))[1],
                                            x0, bounds=bounds, maxiter=maxiter, pass_batch_indices=True)
    assert np.array_equal(np.asarray(xr), xo) and np.array_equal(np.asarray(fr).reshape(-1), fo)
    for a, b in zip(rr, ro):
        assert a.nit == b.nit and a.nfev == b.nfev and a.status == b.status and a.success == b.success
        assert a.message == b.message
    if driver == "direct":  # same rounds, same active sets, same evaluation points
        assert len(trace_ref) == len(trace_ours)
        assert all(np.array_equal(a, b) for a, b in zip(trace_ref, trace_ours))
    monkeypatch.setattr(ours, "_DIRECT", None)


def test_initialize_q_batch_nonneg_matches_reference_algorithm():
    from botorch_b200.optim import initialize_q_batch_nonneg

    X = torch.rand(40, 2, 3, dtype=torch.float64)
    vals = torch.rand(40, dtype=torch.float64)
    vals[::3] = 0.0
    torch.manual_seed(5)
    Xi, vi = initialize_q_batch_nonneg(X, vals, n=6, eta=2.0, alpha=1e-2)
    # restatement of reference initializers.py:1105-1121 with the same RNG stream
    torch.manual_seed(5)
    max_val, max_idx = vals.max(dim=0)
    keep = vals >= 1e-2 * max_val
    idcs = torch.arange(40)[keep][torch.multinomial(torch.exp(2.0 * (vals[keep] / max_val - 1)), 6)]
    if max_idx not in idcs:
        idcs[-1] = max_idx
    assert torch.equal(Xi, X[idcs]) and torch.equal(vi, vals[idcs])
    assert vals.argmax() in idcs
    # fewer positive values than requested: all positives plus random fill
    v2 = torch.zeros(40, dtype=torch.float64)
    v2[[3, 7]] = 1.0
    Xi, vi = initialize_q_batch_nonneg(X, v2, n=5)
    assert Xi.shape[0] == 5 and int((vi > 0).sum()) == 2
    with pytest.warns(Warning, match="nonpositive"):
        initialize_q_batch_nonneg(X, torch.zeros(40, dtype=torch.float64), n=5)


def test_gen_batch_initial_conditions_dispatches_nonneg_and_sample_around_best():
    from botorch_b200.optim import gen_batch_initial_conditions

    c = torch.full((1, 3), 0.3, dtype=torch.float64)
    acqf = QuadraticAcqf(c)
    bounds = torch.stack([torch.zeros(3, dtype=torch.float64), torch.ones(3, dtype=torch.float64)])

    class Shifted(QuadraticAcqf):  # non-negative variant, with a baseline set so that sample_around_best has points
        X_baseline = torch.tensor([[0.31, 0.29, 0.3], [0.9, 0.9, 0.9]], dtype=torch.float64)

        def forward(self, X):
            return (1.0 + super().forward(X)).clamp_min(0.0)

    class _M:
        def posterior(self, X):
            return types.SimpleNamespace(mean=-((X - 0.3) ** 2).sum(-1, keepdim=True))

    sh = Shifted(c)
    sh.model = _M()
    torch.manual_seed(0)
    ics = gen_batch_initial_conditions(sh, bounds, q=1, num_restarts=4, raw_samples=64,
                                       options={"nonnegative": True, "seed": 3, "sample_around_best": True,
                                                "sample_around_best_sigma": 1e-2})
    assert ics.shape == (4, 1, 3)
    # half of the 128 candidates sit within ~3 sigma of a baseline point: the best one must be among the picks
    assert float(((ics - c) ** 2).sum(-1).min()) < 1e-2


def test_retry_on_optimization_warning_and_post_processing_before_argmax():
    from botorch_b200.exceptions.warnings import OptimizationWarning
    from botorch_b200.optim import optimize_acqf

    c = torch.full((1, 2), 0.37, dtype=torch.float64)
    acqf = QuadraticAcqf(c)
    bounds = torch.stack([torch.zeros(2, dtype=torch.float64), torch.ones(2, dtype=torch.float64)])
    calls = {"n": 0}

    def flaky(ics, acq_function, lower_bounds=None, upper_bounds=None, options=None, timeout_sec=None):
        calls["n"] += 1
        if calls["n"] == 1:
            warnings.warn("Optimization failed within `scipy.optimize.minimize` with status 2", OptimizationWarning)
        return ics, acq_function(ics)

    with pytest.warns(RuntimeWarning, match="Trying again with a new set of initial conditions"):
        optimize_acqf(acqf, bounds, q=1, num_restarts=3, raw_samples=16, options={"seed": 0}, gen_candidates=flaky)
    assert calls["n"] == 2
    calls["n"] = 0
    with warnings.catch_warnings(record=True) as ws:
        warnings.simplefilter("always")
        optimize_acqf(acqf, bounds, q=1, num_restarts=3, raw_samples=16, options={"seed": 0}, gen_candidates=flaky,
                      retry_on_optimization_warning=False)
    assert calls["n"] == 1 and any(issubclass(w.category, OptimizationWarning) for w in ws)

    # post-processing (rounding to a 0.25 grid) is applied to ALL restarts and re-evaluated before the arg-max
    ics = torch.tensor([[[0.36, 0.36]], [[0.20, 0.26]]], dtype=torch.float64)
    passthrough = lambda ics, acq_function, **kw: (ics, acq_function(ics))  # noqa: E731
    cand, val = optimize_acqf(acqf, bounds, q=1, num_restarts=2, batch_initial_conditions=ics, gen_candidates=passthrough,
                              post_processing_func=lambda X: (X * 4).round() / 4)
    # restart 0 is best before rounding (0.36 ~ 0.37) but rounds to 0.25; restart 1 rounds to (0.25, 0.25) as well ->
    # equal values, arg-max takes the first; the returned value is the re-evaluated one, a scalar
    assert torch.equal(cand, torch.tensor([[0.25, 0.25]], dtype=torch.float64))
    assert val.ndim == 0 and float(val) == float(acqf(cand.unsqueeze(0)))
    with pytest.raises(Exception, match="not supported for sequential"):
        optimize_acqf(acqf, bounds, q=2, num_restarts=2, batch_initial_conditions=ics, sequential=True)


def test_timeout_stops_the_batched_driver_and_reports_like_the_reference():
    from botorch_b200.optim.batched_lbfgs_b import fmin_l_bfgs_b_batched

    func, x0, bounds = _problem(N=4)
    xs, fs, res = fmin_l_bfgs_b_batched(func, x0, bounds=bounds, maxiter=500, pass_batch_indices=True, timeout_sec=0.0)
    assert all((not r.success) and r.status == 1 and "Optimization timed out after" in r.message for r in res)
    assert all(r.nit == 1 for r in res)  # stopped at the first completed iteration (the reference checks in the callback)
    # the iterate returned is feasible and no worse than the start
    f0, _ = func(np.clip(x0, -1.0, 1.5), list(range(4)))
    assert np.all(fs <= f0 + 1e-12)


def test_scipy_messages_and_iteration_limit_is_not_a_warning():
    from botorch_b200.exceptions.warnings import OptimizationWarning
    from botorch_b200.generation.gen import gen_candidates_scipy

    c = torch.full((1, 4), 0.4, dtype=torch.float64)

    class Rosen(torch.nn.Module):
        def forward(self, X):
            x = X.squeeze(-2)
            return -(100 * (x[..., 1:] - x[..., :-1] ** 2) ** 2 + (1 - x[..., :-1]) ** 2).sum(-1)

    ics = torch.rand(3, 1, 4, dtype=torch.float64)
    with warnings.catch_warnings(record=True) as ws:
        warnings.simplefilter("always")
        gen_candidates_scipy(ics, Rosen(), lower_bounds=0.0, upper_bounds=1.0, options={"maxiter": 2})
    assert not any(issubclass(w.category, OptimizationWarning) for w in ws)
