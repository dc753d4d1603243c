"""CUDA `fused_log_areas` op: the B200 counterpart of `_FusedLogAreas` (reference:
botorch/acquisition/multi_objective/logei.py:107-169, kernel botorch/csrc/logei_fused.cpp:184-374).

`fused_log_areas(obj_subsets, cell_lower, cell_upper, tau_relu, tau_max)` maps
`obj_subsets (B, n_sub, i, m)` and cell bounds `(num_cells, m)` or `(B, num_cells, m)` to the log-areas
`(B, num_cells, n_sub)` of the qLogEHVI / qLogNEHVI inclusion-exclusion loop, differentiable w.r.t.
`obj_subsets`.  The reference only has this kernel on the CPU (gate :320-328).
"""
from __future__ import annotations

import torch
from torch import Tensor

from ... import _lib

_DTYPE = {torch.float64: 0, torch.float32: 1}


def _prep(obj: Tensor, cl: Tensor, cu: Tensor):
    if obj.dim() != 4:
        raise ValueError("obj_subsets must be 4-D")
    if cl.dim() not in (2, 3) or cl.dim() != cu.dim():
        raise ValueError("cell bounds must be 2-D or 3-D")
    if obj.dtype not in _DTYPE:
        raise _lib.McacqError(f"fused_log_areas supports float32/float64, got {obj.dtype}")
    B, n_sub, isz, m = obj.shape
    batched = cl.dim() == 3
    if batched and cl.shape[0] != B:
        raise ValueError("batched cell bounds must have leading dimension B")
    if isz > 32 or m > 8:
        raise ValueError("subset_size or m too large")
    nc = cl.shape[-2]
    return B, n_sub, isz, m, int(batched), nc


class _FusedLogAreasCUDA(torch.autograd.Function):
    @staticmethod
    def forward(ctx, obj_subsets: Tensor, cell_lower: Tensor, cell_upper: Tensor, tau_relu: float, tau_max: float):
        obj = obj_subsets.detach().contiguous()
        cl = cell_lower.detach().to(obj).contiguous()
        cu = cell_upper.detach().to(obj).contiguous()
        _lib.require_cuda(obj, "obj_subsets")
        B, n_sub, isz, m, batched, nc = _prep(obj, cl, cu)
        out = torch.empty(B, nc, n_sub, dtype=obj.dtype, device=obj.device)
        lcl = torch.empty_like(cl)
        rc = _lib.lib().mcacq_log_areas_forward(obj.data_ptr(), cl.data_ptr(), cu.data_ptr(), B, n_sub, isz, m, batched,
                                                nc, _DTYPE[obj.dtype], float(tau_relu), float(tau_max), out.data_ptr(),
                                                lcl.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "mcacq_log_areas_forward")
        ctx.save_for_backward(obj, cl, cu)
        ctx.taus = (float(tau_relu), float(tau_max))
        return out

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        obj, cl, cu = ctx.saved_tensors
        B, n_sub, isz, m, batched, nc = _prep(obj, cl, cu)
        go = grad_out.to(obj).contiguous()
        g_obj = torch.empty_like(obj)
        lcl = torch.empty_like(cl)
        rc = _lib.lib().mcacq_log_areas_backward(go.data_ptr(), obj.data_ptr(), cl.data_ptr(), cu.data_ptr(), B, n_sub,
                                                 isz, m, batched, nc, _DTYPE[obj.dtype], ctx.taus[0], ctx.taus[1],
                                                 g_obj.data_ptr(), lcl.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "mcacq_log_areas_backward")
        return g_obj, None, None, None, None


def fused_log_areas(obj_subsets: Tensor, cell_lower: Tensor, cell_upper: Tensor, tau_relu: float, tau_max: float) -> Tensor:
    return _FusedLogAreasCUDA.apply(obj_subsets, cell_lower, cell_upper, tau_relu, tau_max)


class _FusedLogHVI(torch.autograd.Function):
    """Per-sample log hypervolume improvement in ONE launch (csrc/log_hvi.cu): obj (B, q, m) -> (B,)."""

    @staticmethod
    def forward(ctx, obj: Tensor, cell_lower: Tensor, cell_upper: Tensor, tau_relu: float, tau_max: float):
        o = obj.detach().contiguous()
        cl = cell_lower.detach().to(o).contiguous()
        cu = cell_upper.detach().to(o).contiguous()
        _lib.require_cuda(o, "obj")
        B, q, m = o.shape
        out = torch.empty(B, dtype=torch.float64, device=o.device)
        lcl = torch.empty_like(cl)
        rc = _lib.lib().mcacq_log_hvi_forward(o.data_ptr(), cl.data_ptr(), cu.data_ptr(), B, q, m, cl.shape[-2], float(tau_relu),
                                              float(tau_max), out.data_ptr(), lcl.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "mcacq_log_hvi_forward")
        ctx.save_for_backward(o, cl, cu, out)
        ctx.taus = (float(tau_relu), float(tau_max))
        return out

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        o, cl, cu, out = ctx.saved_tensors
        B, q, m = o.shape
        go = grad_out.to(o).contiguous()
        g_obj = torch.empty_like(o)
        lcl = torch.empty_like(cl)
        rc = _lib.lib().mcacq_log_hvi_backward(go.data_ptr(), out.data_ptr(), o.data_ptr(), cl.data_ptr(), cu.data_ptr(), B, q, m,
                                               cl.shape[-2], ctx.taus[0], ctx.taus[1], g_obj.data_ptr(), lcl.data_ptr(),
                                               _lib.stream_ptr())
        _lib.check(rc, "mcacq_log_hvi_backward")
        return g_obj, None, None, None, None


def log_hvi_fusable(obj: Tensor, cell_lower: Tensor) -> bool:
    return (obj.is_cuda and obj.dtype == torch.float64 and obj.dim() == 3 and obj.shape[-2] <= 6 and obj.shape[-1] <= 4
            and cell_lower.dim() == 2)


def fused_log_hvi(obj: Tensor, cell_lower: Tensor, cell_upper: Tensor, tau_relu: float, tau_max: float) -> Tensor:
    """obj (B, q, m) -> per-sample log HVI (B,), differentiable w.r.t. obj."""
    return _FusedLogHVI.apply(obj, cell_lower, cell_upper, tau_relu, tau_max)
