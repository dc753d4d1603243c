"""Batched L-BFGS-B: many independent bound-constrained problems, ONE batched function call per round.

Role of botorch/optim/batched_lbfgs_b.py (`fmin_l_bfgs_b_batched` :169-362): the reference keeps one
`_lbfgsb.setulb` state machine per restart and evaluates all still-active restarts together.  This
implementation gets the same per-problem iterates without touching scipy's private module: each problem
runs scipy's public `minimize(method="L-BFGS-B", jac=True)` in its own thread, and the threads rendezvous at
every function evaluation so that the coordinator evaluates all pending points with a single call to
`func(X_pending, batch_indices=...)` -- i.e. one fused CUDA forward+backward per round on the GPU.  The
per-problem trajectories are exactly scipy's (same defaults as the reference: m=10, factr via ftol, pgtol
1e-5, maxls 20, maxiter as given); the active set shrinks as restarts converge, so the kernels see any b' >= 1.
"""
from __future__ import annotations

import threading
from typing import Any, Callable

import numpy as np
from scipy.optimize import Bounds, OptimizeResult, minimize


class _Rendezvous:
    def __init__(self, n: int) -> None:
        self.cv = threading.Condition()
        self.pending: dict[int, np.ndarray] = {}
        self.results: dict[int, tuple[float, np.ndarray]] = {}
        self.alive = n
        self.error: BaseException | None = None


def fmin_l_bfgs_b_batched(func: Callable, x0: np.ndarray, bounds=None, maxiter: int = 15000, maxcor: int = 10,
                          ftol: float = 2.2204460492503131e-09, pgtol: float = 1e-5, maxls: int = 20,
                          maxfun: int = 15000, callback: Callable | None = None, pass_batch_indices: bool = False,
                          **unused: Any) -> tuple[np.ndarray, np.ndarray, list[OptimizeResult]]:
    """Minimise N problems `x0[i]` (N x D) sharing `func(X: K x D[, batch_indices]) -> (f: K, g: K x D)`.

    `bounds`: None, a (D x 2) array / list of pairs shared by all problems, or an (N x D x 2) array.
    Returns (xs N x D, fs N, list of scipy OptimizeResult).
    """
    x0 = np.asarray(x0, dtype=np.float64)
    if x0.ndim != 2:
        raise ValueError("x0 must be two-dimensional: (num_problems, dim)")
    N, D = x0.shape
    if bounds is not None:
        barr = np.array([[(-np.inf if lo is None else lo), (np.inf if hi is None else hi)] for lo, hi in bounds],
                        dtype=np.float64) if not isinstance(bounds, np.ndarray) else bounds.astype(np.float64)
        per_problem = barr.ndim == 3
    rv = _Rendezvous(N)
    results: list[OptimizeResult | None] = [None] * N

    def make_fun(i: int):
        def fun(x: np.ndarray):
            with rv.cv:
                rv.pending[i] = np.array(x, dtype=np.float64, copy=True)
                rv.cv.notify_all()
                while i not in rv.results and rv.error is None:
                    rv.cv.wait()
                if rv.error is not None:
                    raise RuntimeError("batched evaluation failed") from rv.error
                return rv.results.pop(i)
        return fun

    def worker(i: int) -> None:
        try:
            b = None
            if bounds is not None:
                bi = barr[i] if per_problem else barr
                b = Bounds(bi[:, 0], bi[:, 1])
            results[i] = minimize(make_fun(i), x0[i], jac=True, method="L-BFGS-B", bounds=b, callback=callback,
                                  options={"maxiter": maxiter, "maxcor": maxcor, "ftol": ftol, "gtol": pgtol,
                                           "maxls": maxls, "maxfun": maxfun})
        except BaseException as e:  # noqa: BLE001 -- surfaced by the coordinator
            with rv.cv:
                if rv.error is None:
                    rv.error = e
        finally:
            with rv.cv:
                rv.alive -= 1
                rv.cv.notify_all()

    threads = [threading.Thread(target=worker, args=(i,), daemon=True) for i in range(N)]
    for t in threads:
        t.start()
    while True:
        with rv.cv:
            while rv.error is None and rv.alive > 0 and len(rv.pending) < rv.alive:
                rv.cv.wait()
            if rv.error is not None or rv.alive == 0:
                break
            idx = sorted(rv.pending)
            X = np.stack([rv.pending.pop(i) for i in idx])
        try:
            f, g = func(X, batch_indices=idx) if pass_batch_indices else func(X)
            f = np.asarray(f, dtype=np.float64).reshape(len(idx))
            g = np.asarray(g, dtype=np.float64).reshape(len(idx), D)
        except BaseException as e:  # noqa: BLE001
            with rv.cv:
                rv.error = e
                rv.cv.notify_all()
            break
        with rv.cv:
            for k, i in enumerate(idx):
                rv.results[i] = (float(f[k]), g[k].copy())
            rv.cv.notify_all()
    for t in threads:
        t.join()
    if rv.error is not None:
        raise rv.error
    xs = np.stack([r.x for r in results])
    fs = np.array([r.fun for r in results])
    return xs, fs, results
