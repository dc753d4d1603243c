"""Package-level flags (role of botorch/settings.py:17-96: small context-manager switches).

`contraction` selects how the dominant contraction `K(X, X_train) @ L^{-T}` (and its transpose in the backward pass)
is executed:
  * "dmma" -- hand-written FP64 tensor-core kernel (`csrc/dgemm_tri.cu`, DMMA.8x8x4);
  * "int8" -- Ozaki-style error-free split onto the INT8 tensor cores (`csrc/ozaki_imma.cu`, tcgen05 + TMEM + TMA).
              The number of signed 8-bit slices (6 or 7 forward, 5..7 backward) is chosen PER FITTED MODEL by a build-time
              probe around the training points so that the posterior variance stays within 1e-9 and gradients within
              1e-7 of the FP64 contraction; a model that no slice count serves runs on "dmma"
              (`DevicePredictionStrategy._select_int8`).  ~2x faster than "dmma".  This is the default.
`int8_slices` = (g_fwd, g_bwd) pins the slice counts and skips the probe (benchmarks / developer tools only).
`int8_cond_limit`: q-batches whose conditioning byte floor(-4 log2 rho) (rho = smallest relative Cholesky pivot of the
              q-batch's conditional covariance) exceeds this limit are re-evaluated through the FP64 contraction; "auto" = the
              library's calibrated limit (`DevicePredictionStrategy.INT8_COND_LIMIT`), None = never re-route.

`int8_max_slices`: evaluate with the most accurate int8 slice counts (7 / 7) instead of the per-model ladder choice; set by the
              L-BFGS-B drivers around their (small, latency-bound) evaluations.

`optimizer` selects what `optimize_acqf` uses when no `gen_candidates` is passed:
  * "scipy"  -- `gen_candidates_scipy`: scipy's own L-BFGS-B routine stepped on the host (iterates bit-identical to
                `scipy.optimize.minimize` per restart, as in the reference), one fused forward+backward per round;
  * "device" -- `gen_candidates_device`: the same algorithm with the state machines resident on the GPU and one CUDA graph
                replay per round (no host hop); candidates agree with "scipy" to optimiser tolerance, not bit for bit.
"""
from __future__ import annotations


class _Flag:
    def __init__(self, default):
        self._value = default

    def value(self):
        return self._value

    def __call__(self, value):
        flag = self

        class _Ctx:
            def __enter__(self_inner):
                self_inner.prev = flag._value
                flag._value = value
                return flag

            def __exit__(self_inner, *exc):
                flag._value = self_inner.prev
                return False

        return _Ctx()

    def set(self, value) -> None:
        self._value = value


contraction = _Flag("int8")
int8_slices = _Flag(None)
int8_cond_limit = _Flag("auto")
int8_max_slices = _Flag(False)
int8_gram = _Flag(True)       # joint posteriors over >= 1024 points: Gram A A^T on the INT8 tensor cores (False: DMMA SYRK)
fused_log_hvi = _Flag(True)   # qLogEHVI: one fused kernel for the inclusion-exclusion loop (False: per-subset-size kernels)
optimizer = _Flag("scipy")
