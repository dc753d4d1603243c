"""torch.autograd bridge to the fused CUDA forward / backward of the MC acquisition value.

Plays the role of `_FusedLogAreas(torch.autograd.Function)` in the reference
(botorch/acquisition/multi_objective/logei.py:107-169): a thin custom op whose forward and backward
call the native library, here through the C ABI of include/mcacq_b200.h.
"""
from __future__ import annotations

import ctypes as C
import warnings
from dataclasses import dataclass

import torch
from torch import Tensor

from .. import _lib
from ..exceptions.errors import NanError, NotPSDError
from ..exceptions.warnings import NumericalWarning
from ..models.prediction_strategy import DevicePredictionStrategy


@dataclass
class BaselineOperands:
    """Device operands of `mcacq_baseline` (qLogNEI `_init_baseline`, acquisition/logei.py:393-459)."""

    U_base: Tensor  # r x d
    A_base: Tensor  # r x np
    L_base: Tensor  # r x r
    desc: _lib.Baseline = None
    A_base_absmax: Tensor = None

    def __post_init__(self):
        r = self.U_base.shape[0]
        self.A_base_absmax = self.A_base.abs().amax(dim=1).contiguous()
        self.desc = _lib.Baseline(r=r, _pad=0, U_base=self.U_base.data_ptr(), A_base=self.A_base.data_ptr(),
                                  L_base=self.L_base.data_ptr(), A_base_absmax=self.A_base_absmax.data_ptr())

    @property
    def r(self) -> int:
        return self.U_base.shape[0]


@dataclass
class MCOperands:
    """Device operands of `mcacq_mc`: transposed base samples and per-sample incumbents."""

    Zt: Tensor  # (r + q) x S
    best: Tensor  # S
    tau_relu: float
    tau_max: float
    fat: bool
    desc: _lib.MC = None
    mode: int | None = None  # utility mode override (2 qEI/qNEI, 3 qSimpleRegret, 4 qPI); None -> int(fat)

    def __post_init__(self):
        self.desc = _lib.MC(S=self.Zt.shape[1], fat=int(self.fat) if self.mode is None else int(self.mode),
                            tau_relu=float(self.tau_relu),
                            tau_max=float(self.tau_max), Zt=self.Zt.data_ptr(), best=self.best.data_ptr())


class LaunchStats:
    """Counts kernels launched through the C ABI (bench.py reports it as `gpu_launches`)."""

    launches = 0

    @classmethod
    def add(cls):
        cls.launches += int(_lib.lib().mcacq_last_launch_count())


def _raise_on_info(info: Tensor) -> None:
    """Mirror the reference's numerical signalling: NumericalWarning on jitter (psd_safe_cholesky),
    NotPSDError after 6 tries, NanError on non-finite samples (utils/low_rank.py:162-171)."""
    flags = int(info.max().item()) if info.numel() else 0  # all status words are non-negative bit sets: max == 0 <=> all clear
    if flags == 0:
        return
    host = info.cpu()
    if bool(((host & _lib.INFO_NOT_PSD) != 0).any()):
        raise NotPSDError("Matrix not positive definite after repeatedly adding jitter up to 1.0e-03.")
    if bool(((host & _lib.INFO_NONFINITE) != 0).any()):
        raise NanError("Samples contain nans or infs.")
    lvl = int((host & _lib.INFO_JITTER_MASK).max())
    if lvl > 0:
        warnings.warn(f"A not p.d., added jitter of {1e-8 * 10 ** (lvl - 1):.1e} to the diagonal", NumericalWarning)


class FusedMCAcquisition(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X: Tensor, strat: DevicePredictionStrategy, base: BaselineOperands | None, mc: MCOperands):
        _lib.require_cuda(X, "X")
        b, q, d = X.shape
        r = base.r if base is not None else 0
        L = _lib.lib()
        Xc = X.detach().contiguous()
        acq = torch.empty(b, device=X.device, dtype=torch.float64)
        info = torch.empty(b, device=X.device, dtype=torch.int32)
        ws = strat.workspace(b, q, r)
        rc = L.mcacq_acq_forward(C.byref(strat.desc), C.byref(base.desc) if base is not None else None,
                                 C.byref(mc.desc), Xc.data_ptr(), b, q, acq.data_ptr(), info.data_ptr(),
                                 ws.data_ptr(), ws.numel(), _lib.stream_ptr())
        _lib.check(rc, "mcacq_acq_forward")
        LaunchStats.add()
        _raise_on_info(info)
        ctx.strat, ctx.base, ctx.mc = strat, base, mc
        ctx.ws = ws
        ctx.save_for_backward(Xc, acq)
        ctx.mark_non_differentiable(info)
        return acq, info

    @staticmethod
    def backward(ctx, grad_acq: Tensor, _grad_info):
        Xc, acq = ctx.saved_tensors
        if ctx.ws is None:
            raise RuntimeError("FusedMCAcquisition backward called twice: the workspace is consumed in place.")
        b, q, d = Xc.shape
        L = _lib.lib()
        g = grad_acq.to(torch.float64).contiguous()
        gX = torch.empty_like(Xc)
        base = ctx.base
        rc = L.mcacq_acq_backward(C.byref(ctx.strat.desc), C.byref(base.desc) if base is not None else None,
                                  C.byref(ctx.mc.desc), Xc.data_ptr(), b, q, acq.data_ptr(), g.data_ptr(),
                                  gX.data_ptr(), ctx.ws.data_ptr(), ctx.ws.numel(), _lib.stream_ptr())
        _lib.check(rc, "mcacq_acq_backward")
        LaunchStats.add()
        ctx.ws = None
        return gX, None, None, None


def fused_acquisition(X: Tensor, strat: DevicePredictionStrategy, base: BaselineOperands | None,
                      mc: MCOperands) -> Tensor:
    """acq[b] for X: b x q x d (fp64, CUDA).  Under `torch.no_grad()` no state is kept."""
    acq, _ = FusedMCAcquisition.apply(X, strat, base, mc)
    return acq
