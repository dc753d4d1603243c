"""CPU oracle for the batched MC-acquisition hot path (TEST INFRASTRUCTURE ONLY).

This package is a gpytorch-free, pure-torch fp64 restatement of the reference's
CPU algorithm for `qLogExpectedImprovement` / `qLogNoisyExpectedImprovement` on an
exact `SingleTaskGP` (BoTorch v0.18.1 + gpytorch>=1.15.2 + linear_operator>=0.6.1).
It exists to CHECK the CUDA path: only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  The product
package `botorch_b200` never imports it and has no CPU fallback.

PARITY STATUS: "parity unpinned" at the gpytorch/linear_operator boundary.  The
arithmetic of kernel evaluation, the train Cholesky, the `L^{-T}` cache, predictive
mean/covariance, `psd_safe_cholesky` and `MultivariateNormal.rsample` lives in
un-vendored third-party packages (pyproject.toml:25-26 of the reference) that are not
installable in the build container, and no reference test pins numbers there.  Those
parts are restated from the published gpytorch 1.15 / linear_operator 0.6 algorithms
(see `oracle/gp.py` docstrings).  The parts that ARE in the reference tree
(`botorch/utils/safe_math.py`, `botorch/sampling/qmc.py`, `botorch/utils/sampling.py`)
are pinned: `tests/test_oracle_golden.py` checks the restatements against the
reference modules imported from `/root/reference` (when present), and
`tests/golden/` holds vectors generated from those reference modules
(`tests/golden/make_golden.py`).  Against an INDEPENDENT implementation of the same
published algorithm the GP part IS pinned: `tests/test_oracle_vs_sklearn.py` compares
posterior mean, full covariance, the train Cholesky factor and `(K + noise)^-1 y` with
scikit-learn's `GaussianProcessRegressor` (fixed hyper-parameters, ARD RBF / Matern-5/2,
ScaleKernel, Normalize, Standardize, homo- / heteroskedastic noise) to 1e-9, and
`tests/test_real_botorch_probe.py` compares the oracle with a real BoTorch + gpytorch
install at 1e-12 wherever one is importable (it skips, visibly, where none is).
"""
