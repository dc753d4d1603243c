"""Package-level flags (role of botorch/settings.py:17-96: small context-manager switches).

`contraction` selects how the dominant contraction `K(X, X_train) @ L^{-T}` (and its transpose in the backward pass)
is executed:
  * "dmma" -- hand-written FP64 tensor-core kernel (`csrc/dgemm_tri.cu`, DMMA.8x8x4);
  * "int8" -- Ozaki-style error-free split onto the INT8 tensor cores (`csrc/ozaki_imma.cu`, tcgen05 + TMEM + TMA),
              6 diagonals of signed 8-bit slices forward / 5 backward: same 1e-9 value and 1e-7 gradient parity, ~2x faster.
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


contraction = _Flag("dmma")
