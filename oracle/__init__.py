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
(`tests/golden/make_golden.py`).
"""
