"""ctypes binding of the C-ABI library `csrc/libmcacq_b200.so` (see include/mcacq_b200.h).

The library is plain C ABI (device pointers + sizes + stream); torch is only used here for device
memory and streams.  There is NO fallback: if the shared library is missing, or CUDA is unavailable
when a compute entry point is called, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import torch

_CSRC = Path(__file__).resolve().parent / "csrc"
_SO = _CSRC / "libmcacq_b200.so"

KERNEL_RBF = 0
KERNEL_MATERN52 = 1
TRI_UPPER, TRI_LOWER, TRI_DENSE = 0, 1, 2
INFO_JITTER_MASK, INFO_NOT_PSD, INFO_NONFINITE = 0x7, 0x8, 0x10
INFO_FLAG_MASK, INFO_COND_SHIFT, INFO_COND_MASK = 0x1F, 8, 0xFF00
INFO_VAR_SHIFT, INFO_VAR_MASK = 16, 0xFF0000
MAX_Q, MAX_D, MAX_R = 32, 64, 512

_ERRORS = {-1: "MCACQ_EINVAL (bad argument)", -2: "MCACQ_ELIMIT (q/r/d/S outside compiled limits)",
           -3: "MCACQ_EWORKSPACE (workspace too small)"}


class McacqError(RuntimeError):
    pass


class Model(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("d", C.c_int32), ("np", C.c_int32), ("kernel_id", C.c_int32),
        ("outputscale", C.c_double), ("mean_const", C.c_double), ("y_mean", C.c_double), ("y_std", C.c_double),
        ("x_offset", C.c_void_p), ("x_coef", C.c_void_p), ("lengthscale", C.c_void_p), ("U_train", C.c_void_p),
        ("alpha", C.c_void_p), ("R", C.c_void_p), ("Rt", C.c_void_p),
        ("contraction", C.c_int32), ("g_fwd", C.c_int32), ("g_bwd", C.c_int32), ("_pad", C.c_int32),
        ("Rt_slices", C.c_void_p), ("Rt_scale", C.c_void_p), ("R_slices", C.c_void_p), ("R_scale", C.c_void_p),
    ]


class Baseline(C.Structure):
    _fields_ = [("r", C.c_int32), ("_pad", C.c_int32), ("U_base", C.c_void_p), ("A_base", C.c_void_p),
                ("L_base", C.c_void_p), ("A_base_absmax", C.c_void_p)]


class MC(C.Structure):
    _fields_ = [("S", C.c_int32), ("fat", C.c_int32), ("tau_relu", C.c_double), ("tau_max", C.c_double),
                ("Zt", C.c_void_p), ("best", C.c_void_p),
                ("obj_weight", C.c_double), ("obj_offset", C.c_double), ("util_param", C.c_double), ("Zbar", C.c_void_p),
                ("n_con", C.c_int32), ("con_fat", C.c_int32),
                ("con_a", C.c_double * 4), ("con_b", C.c_double * 4), ("con_eta", C.c_double * 4),
                ("jitter_f32", C.c_int32), ("_pad", C.c_int32)]


def build(force: bool = False) -> Path:
    """Compile the shared library in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    if force:
        subprocess.run(["make", "-C", str(_CSRC), "clean"], check=True, capture_output=True)
    res = subprocess.run(["make", "-C", str(_CSRC), "-j", str(os.cpu_count() or 4)], capture_output=True, text=True)
    if res.returncode != 0:
        raise McacqError(f"nvcc build of {_SO.name} failed:\n{res.stdout}\n{res.stderr}")
    return _SO


_lib = None

EXPORTS = [
    "mcacq_version", "mcacq_num_sms", "mcacq_scale_inputs", "mcacq_cov_cross", "mcacq_cov_cross_bwd",
    "mcacq_dgemm_tri", "mcacq_workspace_bytes", "mcacq_posterior", "mcacq_posterior_backward", "mcacq_acq_forward", "mcacq_acq_backward",
    "mcacq_last_launch_count", "mcacq_sobol_draw", "mcacq_log_areas_forward", "mcacq_log_areas_backward", "mcacq_log_hvi_forward", "mcacq_log_hvi_backward", "mcacq_dgemm_nt", "mcacq_syrk_sub", "mcacq_lower_times_few", "mcacq_slice_rows",
    "mcacq_ozaki_contract", "mcacq_cov_cross_sliced", "mcacq_workspace_bytes_model",
    "mcacq_sample_reduce_forward", "mcacq_info_summary", "mcacq_lbfgsb_state_bytes", "mcacq_lbfgsb_init", "mcacq_lbfgsb_step", "mcacq_lbfgsb_summary",
    "mcacq_fused_supported", "mcacq_fast_math_probe",
]


def lib() -> C.CDLL:
    """Load (once) and return the C-ABI library; raises loudly if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not _SO.exists():
        raise McacqError(
            f"{_SO} not found: the CUDA extension is not built. Run `python -c 'import __graft_entry__ as g; "
            "g.build()'` (or `make -C botorch_b200/csrc`). There is no CPU fallback."
        )
    L = C.CDLL(str(_SO))
    vp, i32, i64, dbl, sz = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_size_t
    L.mcacq_version.restype = C.c_char_p
    L.mcacq_num_sms.restype = i32
    L.mcacq_last_launch_count.restype = i32
    L.mcacq_sobol_draw.argtypes = [vp, vp, i32, i64, i64, vp, vp]
    L.mcacq_fast_math_probe.argtypes = [i32, vp, vp, i64, vp]
    L.mcacq_fused_supported.argtypes = [i32, i32, i32, i32]
    L.mcacq_scale_inputs.argtypes = [vp, i64, i32, vp, vp, vp, vp, vp]
    L.mcacq_cov_cross.argtypes = [i32, dbl, vp, i64, vp, i32, i32, vp, i64, vp]
    L.mcacq_cov_cross_bwd.argtypes = [i32, dbl, vp, i64, vp, i32, i32, vp, i64, vp, vp, vp, i32, vp]
    L.mcacq_dgemm_tri.argtypes = [i32, i64, i32, vp, vp, vp, vp, vp]
    L.mcacq_dgemm_nt.argtypes = [i32, i64, i32, i32, vp, i64, vp, i64, vp, i64, vp, vp]
    L.mcacq_syrk_sub.argtypes = [i64, i32, vp, i64, vp, i64, dbl, vp, vp]
    L.mcacq_lower_times_few.argtypes = [i64, i32, vp, i64, vp, i64, vp, i64, vp]
    L.mcacq_cov_cross_sliced.argtypes = [i32, dbl, vp, i64, vp, i32, i32, i64, vp, i32, i32, vp, vp, vp]
    L.mcacq_slice_rows.argtypes = [vp, i64, i32, i64, i32, i32, i32, i32, vp, vp, vp]
    L.mcacq_ozaki_contract.argtypes = [i32, i64, i32, i32, i32, vp, vp, vp, vp, vp, i64, vp]
    L.mcacq_workspace_bytes.argtypes = [i64, i32, i32, i32, i32]
    L.mcacq_workspace_bytes.restype = sz
    L.mcacq_workspace_bytes_model.argtypes = [C.POINTER(Model), i64, i32, i32]
    L.mcacq_workspace_bytes_model.restype = sz
    L.mcacq_sample_reduce_forward.argtypes = [C.POINTER(Baseline), C.POINTER(MC), vp, vp, vp, i64, i32, vp, vp, vp, vp, vp]
    L.mcacq_info_summary.argtypes = [vp, i64, vp, vp]
    L.mcacq_log_hvi_forward.argtypes = [vp, vp, vp, i64, i32, i32, i32, dbl, dbl, vp, vp, vp]
    L.mcacq_log_hvi_backward.argtypes = [vp, vp, vp, vp, vp, i64, i32, i32, i32, dbl, dbl, vp, vp, vp]
    L.mcacq_lbfgsb_state_bytes.argtypes = [i64, i32]
    L.mcacq_lbfgsb_state_bytes.restype = sz
    L.mcacq_lbfgsb_init.argtypes = [i64, i32, vp, vp, vp, vp, vp, vp]
    L.mcacq_lbfgsb_step.argtypes = [i64, i32, vp, vp, vp, dbl, vp, vp, dbl, dbl, i32, i32, i32, vp, vp, vp]
    L.mcacq_lbfgsb_summary.argtypes = [i64, i32, vp, dbl, vp, vp, vp]
    L.mcacq_posterior.argtypes = [C.POINTER(Model), vp, i64, i32, vp, vp, vp, sz, vp]
    L.mcacq_posterior_backward.argtypes = [C.POINTER(Model), vp, i64, i32, vp, vp, vp, vp, sz, vp]
    L.mcacq_acq_forward.argtypes = [C.POINTER(Model), C.POINTER(Baseline), C.POINTER(MC), vp, i64, i32, vp, vp, vp, sz, vp]
    L.mcacq_acq_backward.argtypes = [C.POINTER(Model), C.POINTER(Baseline), C.POINTER(MC), vp, i64, i32, vp, vp, vp, vp, sz, vp]
    L.mcacq_log_areas_forward.argtypes = [vp, vp, vp, i64, i32, i32, i32, i32, i32, i32, dbl, dbl, vp, vp, vp]
    L.mcacq_log_areas_backward.argtypes = [vp, vp, vp, vp, i64, i32, i32, i32, i32, i32, i32, dbl, dbl, vp, vp, vp]
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in ("mcacq_version", "mcacq_workspace_bytes", "mcacq_workspace_bytes_model", "mcacq_lbfgsb_state_bytes"):
            fn.restype = i32
    _lib = L
    return L


def fused_supported(q: int, r: int, S: int, mc_mean: bool = False) -> bool:
    """Whether `mcacq_acq_forward / _backward` take this shape (compiled limits + the shared memory of the sample / reduce
    kernels, which grows with q * r); host-only query."""
    return bool(lib().mcacq_fused_supported(int(q), int(r), int(S), int(bool(mc_mean))))


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    if rc < 0:
        raise McacqError(f"{what}: {_ERRORS.get(rc, rc)}")
    raise McacqError(f"{what}: CUDA error {rc} ({torch.cuda.get_device_name() if torch.cuda.is_available() else 'no GPU'})")


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise McacqError(f"{name} must be a CUDA tensor (got {t.device}); botorch_b200 has no CPU path.")
    if t.dtype not in (torch.float64, torch.float32, torch.int32, torch.int8):
        raise McacqError(f"{name} must be float64 (got {t.dtype}).")
    if not t.is_contiguous():
        raise McacqError(f"{name} must be contiguous.")
    if t.device.index is not None and t.device.index != torch.cuda.current_device():
        # the kernels are launched on the CURRENT device's stream (one process per GPU: torch.cuda.set_device(LOCAL_RANK))
        raise McacqError(f"{name} lives on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}; "
                         "call torch.cuda.set_device(...) or wrap the call in `with torch.cuda.device(...)`.")


def ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr() -> int:
    """The current CUDA stream of the current device as a `cudaStream_t` value.  `torch.cuda.current_stream()` builds a Python
    Stream object through several device-index look-ups (~25 us, three to four times per optimiser round); the raw accessor
    returns the same handle in ~1 us."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m
