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

from .. import _lib, settings
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
    mode: int | None = None  # utility mode override (2 qEI/qNEI, 3 qSimpleRegret, 4 qPI, 5 qUCB/qLCB, 6 qPSTD); None -> int(fat)
    obj_weight: float = 1.0  # affine objective obj = w y + o (LinearMCObjective on the single outcome)
    obj_offset: float = 0.0
    util_param: float = 0.0  # modes 5 / 6
    constraints: tuple = ()  # ((a, b, eta), ...): smoothed indicators of a y + b <= 0
    con_fat: bool = False
    Zbar: Tensor = None

    def __post_init__(self):
        mode = int(self.fat) if self.mode is None else int(self.mode)
        if mode >= 5:
            self.Zbar = self.Zt.mean(dim=1).contiguous()
        cons = list(self.constraints)
        arr = lambda k: (C.c_double * 4)(*([c[k] for c in cons] + [1.0 if k == 2 else 0.0] * (4 - len(cons))))  # noqa: E731
        self.desc = _lib.MC(S=self.Zt.shape[1], fat=mode, tau_relu=float(self.tau_relu),
                            tau_max=float(self.tau_max), Zt=self.Zt.data_ptr(), best=self.best.data_ptr(),
                            obj_weight=float(self.obj_weight), obj_offset=float(self.obj_offset),
                            util_param=float(self.util_param), Zbar=_lib.ptr(self.Zbar), n_con=len(cons),
                            con_fat=int(bool(self.con_fat)), con_a=arr(0), con_b=arr(1), con_eta=arr(2),
                            jitter_f32=int(torch.get_default_dtype() == torch.float32), _pad=0)


class LaunchStats:
    """Counts kernels launched through the C ABI (bench.py reports it as `gpu_launches`)."""

    launches = 0

    @classmethod
    def add(cls):
        cls.launches += int(_lib.lib().mcacq_last_launch_count())


def _info_summary(info: Tensor) -> tuple[int, int, int]:
    """(OR of the flag bits, largest conditioning byte, largest variance-collapse byte) of the status words: one small launch
    (`mcacq_info_summary`) and one 12-byte read -- the only device synchronisation of a fused forward call."""
    if not info.numel():
        return 0, 0, 0
    out = torch.empty(3, dtype=torch.int32, device=info.device)
    _lib.check(_lib.lib().mcacq_info_summary(info.data_ptr(), info.numel(), out.data_ptr(), _lib.stream_ptr()), "mcacq_info_summary")
    LaunchStats.launches += 1
    flags, cond, vbyte = out.tolist()
    return int(flags), int(cond), int(vbyte)


def _raise_on_info(info: Tensor, flags: int | None = None) -> None:
    """Mirror the reference's numerical signalling: NumericalWarning on jitter (psd_safe_cholesky),
    NotPSDError after 6 tries, NanError on non-finite samples (utils/low_rank.py:162-171)."""
    if flags is None:
        flags = _info_summary(info)[0]
    if flags == 0:
        return
    host = info.cpu() & _lib.INFO_FLAG_MASK
    if bool(((host & _lib.INFO_NOT_PSD) != 0).any()):
        raise NotPSDError("Matrix not positive definite after repeatedly adding jitter up to 1.0e-03.")
    if bool(((host & _lib.INFO_NONFINITE) != 0).any()):
        raise NanError("Samples contain nans or infs.")
    # one warning per escalation level that occurred, in the order psd_safe_cholesky would have emitted them for the batch
    # (it warns once per retry round: every level up to the largest needed by any element)
    levels = sorted(set((host & _lib.INFO_JITTER_MASK).tolist()) - {0})
    if levels:
        counts = {lv: int(((host & _lib.INFO_JITTER_MASK) == lv).sum()) for lv in levels}
        for lv in range(1, levels[-1] + 1):
            n_here = counts.get(lv, 0)
            warnings.warn(f"A not p.d., added jitter of {1e-8 * 10 ** (lv - 1):.1e} to the diagonal"
                          + (f" ({n_here} of {host.numel()} q-batches became p.d. at this level)" if n_here else ""),
                          NumericalWarning)


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


class _DropRows(torch.autograd.Function):
    """Identity whose backward zeroes the gradient rows of the q-batches that were re-evaluated elsewhere (their int8
    gradient rows are unused, and may be non-finite if the fixed-point blocks happened not to be positive definite)."""

    @staticmethod
    def forward(ctx, X: Tensor, holder: dict):
        ctx.holder = holder
        return X.view_as(X)

    @staticmethod
    def backward(ctx, g: Tensor):
        idx = ctx.holder.get("idx")
        return (g if idx is None else g.index_fill(0, idx, 0.0)), None


class RerouteStats:
    """How many q-batches the int8 mode sent to the FP64 contraction because of their conditioning (diagnostics / tests)."""

    q_batches = 0
    calls = 0


def fused_acquisition(X: Tensor, strat: DevicePredictionStrategy, base: BaselineOperands | None,
                      mc: MCOperands) -> Tensor:
    """acq[b] for X: b x q x d (fp64, CUDA).  Under `torch.no_grad()` no state is kept.

    int8 contraction mode: the fixed-point contraction reproduces the posterior blocks to an ABSOLUTE accuracy (a fraction of
    the prior variance, which `DevicePredictionStrategy._select_int8` bounds per fitted model), and a q-batch whose joint
    covariance is nearly singular amplifies any absolute perturbation by 1 / rho (rho = smallest relative Cholesky pivot,
    reported per q-batch in the status word), and a point whose variance has collapsed to a small fraction of the prior has
    lost that many digits of `prior - |a|^2`.  q-batches beyond `strat.int8_cond_limit` or `strat.int8_var_byte_limit` are
    therefore re-evaluated -- value and gradient -- through the FP64 DMMA contraction, so that the mode's accuracy does not
    depend on where X lies."""
    int8 = strat.contraction == "int8"
    if int8 and settings.int8_max_slices.value():
        strat = strat.max_slices_view()
    limit = strat.int8_cond_limit if int8 else None
    vlimit = strat.int8_var_byte_limit if int8 else None
    holder = {}
    X8 = _DropRows.apply(X, holder) if (int8 and X.requires_grad and torch.is_grad_enabled()) else X
    acq, info = FusedMCAcquisition.apply(X8, strat, base, mc)
    flags, cond, vbyte = _info_summary(info)
    if (limit is not None and cond > limit) or (vlimit is not None and vbyte > vlimit):
        bad = torch.zeros_like(info, dtype=torch.bool)
        if limit is not None:
            bad |= ((info >> _lib.INFO_COND_SHIFT) & 0xFF).gt(limit)
        if vlimit is not None:
            bad |= ((info >> _lib.INFO_VAR_SHIFT) & 0xFF).gt(vlimit)
        idx = bad.nonzero().squeeze(-1)
        holder["idx"] = idx
        acq64, info64 = FusedMCAcquisition.apply(X.index_select(0, idx), strat.fp64_view(), base, mc)
        acq = acq.index_copy(0, idx, acq64)       # the int8 results of these q-batches (and their gradient path) are dropped
        info = info.index_copy(0, idx, info64)
        RerouteStats.q_batches += int(idx.numel())
        RerouteStats.calls += 1
        flags = _info_summary(info)[0]
    _raise_on_info(info, flags)
    return acq


def affine_constraints(constraints, eta, fat, device=None) -> tuple | None:
    """Compile outcome constraints for the fused kernels.  The callables are arbitrary Python (reference contract:
    `sample_shape x batch x q x m` samples -> `sample_shape x batch x q`, feasible iff <= 0), but on a single-outcome
    model the usual ones are affine in the sample value, `c(y) = a y + b`; those -- at most four, detected by probing the
    callable -- are evaluated inside `sample_reduce` (smoothed indicator of botorch/utils/objective.py:135-211).  Returns
    ((a, b, eta), ...) or None when a constraint is not affine / not fusable (the caller then takes the generic route)."""
    if constraints is None:
        return ()
    if len(constraints) > 4 or isinstance(fat, list):
        return None
    etas = eta if isinstance(eta, Tensor) else torch.full((len(constraints),), float(eta))
    if etas.numel() != len(constraints) or not bool((etas > 0).all()):
        return None
    probe = torch.tensor([-1.3, 0.0, 2.1, 5.7, -40.0], dtype=torch.float64, device=device).view(1, 1, 5, 1)
    out = []
    for con, e in zip(constraints, etas.tolist()):
        try:
            c = con(probe)
        except Exception:  # noqa: BLE001 -- e.g. a callable that indexes a second outcome
            return None
        if not isinstance(c, Tensor) or c.shape != probe.shape[:-1]:
            return None
        c = c.reshape(-1).to(torch.float64).cpu()
        y = probe.reshape(-1).cpu()
        b0 = float(c[1])
        a0 = float((c[2] - c[1]) / (y[2] - y[1]))
        if not torch.allclose(c, a0 * y + b0, rtol=1e-12, atol=1e-12):
            return None
        out.append((a0, b0, float(e)))
    return tuple(out)
