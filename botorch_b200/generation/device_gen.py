"""`gen_candidates_device`: multi-start L-BFGS-B with the optimiser state resident on the GPU.

Same contract as `gen_candidates_scipy` (reference: botorch/generation/gen.py:62-507, box-bounded fast path) -- initial
conditions in, `(candidates, acquisition values)` out -- but the per-restart L-BFGS-B state machines that the reference
steps on the host (botorch/optim/batched_lbfgs_b.py:365-634) live in HBM and are advanced by `mcacq_lbfgsb_step`
(csrc/lbfgsb.cu).  One optimiser round is

    fused forward  ->  fused backward  ->  L-BFGS-B step        (all restarts, one CUDA graph replay)

with no host <-> device hop: the step kernel reads the acquisition values and gradients where the backward kernels left
them and writes the next trial points into the buffer the forward kernels read.  The host only polls the number of active
restarts every `check_every` rounds.  Restarts that have converged keep their final iterate (their rows are still
evaluated -- the shapes stay static, which is what makes the round graph-capturable -- but their state no longer moves).

The algorithm is L-BFGS-B 3.0 as scipy implements it; iterates agree with scipy's to rounding on the same function
(oracle/lbfgsb.py is the CPU restatement both are tested against), so candidates differ from `gen_candidates_scipy` only
by how rounding differences propagate through the iterations: candidate parity is tolerance-based, as SURVEY.md section
8f N4 anticipates.  Acquisition functions that are not on the fused route are optimised through the same device state
machines with values and gradients from torch autograd.
"""
from __future__ import annotations

import ctypes as C
import time
import warnings

import torch
from torch import Tensor

from .. import _lib
from ..exceptions.errors import OptimizationGradientError, UnsupportedError
from ..exceptions.warnings import OptimizationWarning
from ..optim.utils import columnwise_clamp

TASK_ACTIVE, TASK_CONVERGED, TASK_STOPPED, TASK_ABNORMAL = 0, 2, 3, 4
_MESSAGES = {401: "CONVERGENCE: NORM OF PROJECTED GRADIENT <= PGTOL", 402: "CONVERGENCE: RELATIVE REDUCTION OF F <= FACTR*EPSMCH",
             502: "STOP: TOTAL NO. OF F,G EVALUATIONS EXCEEDS LIMIT", 504: "STOP: TOTAL NO. OF ITERATIONS REACHED LIMIT",
             800: "ABNORMAL: ", 0: ""}


class DeviceLBFGSB:
    """N independent bound-constrained problems of dimension D, state in one device buffer (`mcacq_lbfgsb_*`)."""

    def __init__(self, x0: Tensor, lower: Tensor, upper: Tensor, maxiter: int = 15000, maxfun: int = 15000,
                 ftol: float = 2.2204460492503131e-09, pgtol: float = 1e-5, maxls: int = 20) -> None:
        _lib.require_cuda(x0, "x0")
        if x0.dim() != 2:
            raise ValueError("x0 must be N x D")
        self.N, self.D = x0.shape
        dev = x0.device
        f64 = dict(device=dev, dtype=torch.float64)
        self.lower = lower.to(**f64).expand(self.D).contiguous()
        self.upper = upper.to(**f64).expand(self.D).contiguous()
        self.factr = float(ftol) / torch.finfo(torch.float64).eps
        self.pgtol, self.maxiter, self.maxfun, self.maxls = float(pgtol), int(maxiter), int(maxfun), int(maxls)
        L = _lib.lib()
        self.X = torch.empty(self.N, self.D, **f64)
        self.state = torch.zeros(L.mcacq_lbfgsb_state_bytes(self.N, self.D), dtype=torch.uint8, device=dev)
        self.n_active = torch.zeros(1, dtype=torch.int32, device=dev)
        self.reset(x0)

    def reset(self, x0: Tensor) -> None:
        """Restart all N problems from `x0` (clamped into the box); buffers and their addresses stay the same."""
        x0c = x0.to(device=self.X.device, dtype=torch.float64).contiguous()
        _lib.check(_lib.lib().mcacq_lbfgsb_init(self.N, self.D, x0c.data_ptr(), self.lower.data_ptr(), self.upper.data_ptr(),
                                                self.X.data_ptr(), self.state.data_ptr(), _lib.stream_ptr()),
                   "mcacq_lbfgsb_init")
        self.n_active.fill_(self.N)

    def step(self, f: Tensor, g: Tensor, sign: float = 1.0) -> None:
        """Feed f [N], g [N x D] evaluated at `self.X`; `self.X` then holds the next trial points."""
        _lib.check(_lib.lib().mcacq_lbfgsb_step(self.N, self.D, self.X.data_ptr(), f.data_ptr(), g.data_ptr(), float(sign),
                                                self.lower.data_ptr(), self.upper.data_ptr(), self.factr, self.pgtol,
                                                self.maxiter, self.maxfun, self.maxls, self.state.data_ptr(),
                                                self.n_active.data_ptr(), _lib.stream_ptr()), "mcacq_lbfgsb_step")

    def active(self) -> int:
        return int(self.n_active.item())

    def summary(self, sign: float = 1.0) -> tuple[Tensor, Tensor]:
        """(f [N] with the sign removed, status [N x 4] = task, message, iterations, evaluations) on the host."""
        f = torch.empty(self.N, device=self.X.device, dtype=torch.float64)
        status = torch.empty(self.N, 4, device=self.X.device, dtype=torch.int32)
        _lib.check(_lib.lib().mcacq_lbfgsb_summary(self.N, self.D, self.state.data_ptr(), float(sign), f.data_ptr(),
                                                   status.data_ptr(), _lib.stream_ptr()), "mcacq_lbfgsb_summary")
        return f.cpu(), status.cpu()


class _FusedRound:
    """forward + backward + step over static buffers, captured once as a CUDA graph and replayed every round."""

    def __init__(self, acqf, opt: DeviceLBFGSB, q: int, d: int, use_graph: bool = True) -> None:
        from ..acquisition._fused import LaunchStats

        self.opt, self.q, self.d = opt, q, d
        N = opt.N
        # no host between the launches of a captured round, hence no re-routing of ill-conditioned / collapsed q-batches: the
        # rounds use the int8 mode's most accurate slice counts instead (FP64-level; a round is a few hundred rows)
        self.strat = acqf.model.prediction_strategy().max_slices_view()
        self.base = acqf._baseline_operands() if hasattr(acqf, "_baseline_operands") else None
        Xv = opt.X.view(N, q, d)
        self.mc = acqf._mc_operands(Xv)
        r = self.base.r if self.base is not None else 0
        dev = opt.X.device
        f64 = dict(device=dev, dtype=torch.float64)
        self.acq = torch.empty(N, **f64)
        self.info = torch.zeros(N, dtype=torch.int32, device=dev)
        self.info_or = torch.zeros(N, dtype=torch.int32, device=dev)   # sticky: flags of every round (checked once, lazily)
        self.gX = torch.empty(N, q, d, **f64)
        self.ones = torch.ones(N, **f64)
        self.ws = self.strat.workspace(N, q, r)
        self.stats = LaunchStats
        self.graph = None
        self.launches_per_round = 0
        self.use_graph = use_graph
        self._launch()             # warm-up outside capture (first-call attribute / occupancy queries, tensor-map encoder)
        self.rounds = 1
        if use_graph:
            try:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._launch()
                self.graph = g
                self.rounds += 1   # the capture pass does not execute; the first replay below is round 2
            except Exception:  # noqa: BLE001 -- capture is an optimisation; eager launches are always valid
                self.graph = None
                torch.cuda.synchronize()

    def _launch(self) -> None:
        L, st = _lib.lib(), _lib.stream_ptr()
        o, N, q = self.opt, self.opt.N, self.q
        base = C.byref(self.base.desc) if self.base is not None else None
        _lib.check(L.mcacq_acq_forward(C.byref(self.strat.desc), base, C.byref(self.mc.desc), o.X.data_ptr(), N, q,
                                       self.acq.data_ptr(), self.info.data_ptr(), self.ws.data_ptr(), self.ws.numel(), st),
                   "mcacq_acq_forward")
        n1 = int(L.mcacq_last_launch_count())
        _lib.check(L.mcacq_acq_backward(C.byref(self.strat.desc), base, C.byref(self.mc.desc), o.X.data_ptr(), N, q,
                                        self.acq.data_ptr(), self.ones.data_ptr(), self.gX.data_ptr(), self.ws.data_ptr(),
                                        self.ws.numel(), st), "mcacq_acq_backward")
        n2 = int(L.mcacq_last_launch_count())
        self.info_or.bitwise_or_(self.info)
        o.step(self.acq, self.gX.view(N, -1), sign=-1.0)
        self.launches_per_round = n1 + n2 + 2

    def restart(self, x0: Tensor) -> None:
        """Re-use the captured round for a new set of initial conditions (same shapes, same operands, same options)."""
        self.opt.reset(x0)
        self.info_or.zero_()
        self.rounds = 0

    def run(self) -> None:
        if self.graph is not None:
            self.graph.replay()
        else:
            self._launch()
        self.rounds += 1
        self.stats.launches += self.launches_per_round


def _cached_round(acqf, opt: DeviceLBFGSB, clamped: Tensor, q: int, d: int, options: dict) -> _FusedRound:
    """Capturing a round costs milliseconds, an optimiser round 0.1-1 ms: the captured graph (with its static buffers and
    optimiser state) is kept on the acquisition function and re-used by later calls with the same shapes, operands and
    L-BFGS-B options (sequential greedy optimisation, retries, repeated `optimize_acqf` calls)."""
    use_graph = bool(options.get("cuda_graph", True))
    strat = acqf.model.prediction_strategy().max_slices_view()   # (what _FusedRound runs; cached on the strategy)
    base = acqf._baseline_operands() if hasattr(acqf, "_baseline_operands") else None
    mc = acqf._mc_operands(opt.X.view(opt.N, q, d))
    key = (opt.N, q, d, opt.maxiter, opt.maxfun, opt.factr, opt.pgtol, opt.maxls, use_graph, id(strat), id(base), id(mc),
           tuple(opt.lower.tolist()), tuple(opt.upper.tolist()))
    cache = acqf.__dict__.setdefault("_device_rounds", {})
    rnd = cache.get(key)
    if rnd is not None and rnd.strat is strat and rnd.base is base and rnd.mc is mc:
        rnd.restart(clamped.reshape(opt.N, q * d))
        return rnd
    cache.clear()   # one live graph per acquisition function (each holds a workspace)
    rnd = cache[key] = _FusedRound(acqf, opt, q, d, use_graph=use_graph)
    return rnd


def _fused_capable(acqf, X: Tensor) -> bool:
    return (hasattr(acqf, "_fusable") and hasattr(acqf, "_mc_operands") and acqf._fusable(X)
            and getattr(acqf, "X_pending", None) is None
            and (not hasattr(acqf, "_baseline_operands")
                 or (getattr(acqf, "_cache_root", False) and hasattr(acqf, "_baseline_L")
                     and acqf.X_baseline.dim() == 2
                     and _lib.fused_supported(X.shape[-2], acqf.X_baseline.shape[-2], acqf.sample_shape.numel())))
            and (not hasattr(acqf, "best_f") or acqf.best_f.numel() == 1))


def gen_candidates_device(initial_conditions: Tensor, acquisition_function, lower_bounds=None, upper_bounds=None,
                          inequality_constraints=None, equality_constraints=None, nonlinear_inequality_constraints=None,
                          options: dict | None = None, fixed_features=None, timeout_sec: float | None = None,
                          **unused) -> tuple[Tensor, Tensor]:
    if inequality_constraints or equality_constraints or nonlinear_inequality_constraints or fixed_features:
        raise UnsupportedError("botorch_b200.gen_candidates_device implements the box-bounded L-BFGS-B path only.")
    options = dict(options or {})
    options.setdefault("maxiter", 2000)
    if options.get("method", "L-BFGS-B") != "L-BFGS-B" or not options.get("with_grad", True):
        raise UnsupportedError("Only method='L-BFGS-B' with gradients is supported.")
    if not initial_conditions.is_cuda:
        raise _lib.McacqError("gen_candidates_device needs CUDA initial conditions; botorch_b200 has no CPU path.")
    orig_shape = initial_conditions.shape
    if initial_conditions.ndim == 2:
        initial_conditions = initial_conditions.unsqueeze(0)
    clamped = columnwise_clamp(X=initial_conditions, lower=lower_bounds, upper=upper_bounds, raise_on_violation=True)
    nb, q, d = clamped.shape
    dev = clamped.device
    f64 = dict(device=dev, dtype=torch.float64)
    inf = float("inf")
    lo = torch.full((d,), -inf, **f64) if lower_bounds is None else torch.as_tensor(lower_bounds, **f64).expand(d)
    hi = torch.full((d,), inf, **f64) if upper_bounds is None else torch.as_tensor(upper_bounds, **f64).expand(d)
    opt = DeviceLBFGSB(clamped.reshape(nb, q * d).to(torch.float64), lo.repeat(q), hi.repeat(q),
                       maxiter=int(options["maxiter"]), maxfun=int(options.get("maxfun", 15000)),
                       ftol=float(options.get("ftol", 2.2204460492503131e-09)), pgtol=float(options.get("pgtol", 1e-5)),
                       maxls=int(options.get("maxls", 20)))
    check_every = int(options.get("check_every", 8))
    max_rounds = int(options.get("maxfun", 15000)) + 1
    start = time.monotonic()
    timed_out = False
    Xv = opt.X.view(nb, q, d)
    if _fused_capable(acquisition_function, Xv):
        rnd = _cached_round(acquisition_function, opt, clamped, q, d, options)
        opt = rnd.opt
        done = rnd.rounds
        while done < max_rounds:
            rnd.run()
            done += 1
            if done % check_every == 0:
                if opt.active() == 0:
                    break
                if timeout_sec is not None and time.monotonic() - start > timeout_sec:
                    timed_out = True
                    break
        info = rnd.info_or
        from ..acquisition._fused import _raise_on_info

        _raise_on_info(info)  # one lazy check for the whole optimisation (jitter warnings, NotPSD / NaN errors)
    else:
        done = 0
        while done < max_rounds:
            X = Xv.detach().clone().requires_grad_(True)
            acq = acquisition_function(X)
            (gX,) = torch.autograd.grad(acq.sum(), X)
            opt.step(acq.detach().to(torch.float64).contiguous(), gX.reshape(nb, -1).to(torch.float64).contiguous(), sign=-1.0)
            done += 1
            if done % check_every == 0:
                if opt.active() == 0:
                    break
                if timeout_sec is not None and time.monotonic() - start > timeout_sec:
                    timed_out = True
                    break
    _, status = opt.summary(sign=-1.0)
    if torch.isnan(opt.X).any():
        raise OptimizationGradientError("NaN in the optimiser iterates: the acquisition gradient contained NaNs. This often "
                                        "indicates numerical issues.", current_x=opt.X.detach().cpu().numpy())
    import logging

    logger = logging.getLogger("botorch")
    for task, msg, nit, nfev in status.tolist():
        if task == TASK_ABNORMAL:
            with warnings.catch_warnings():
                warnings.simplefilter("always", category=OptimizationWarning)
                warnings.warn("Optimization failed within `scipy.optimize.minimize` with status 2 and message "
                              f"{_MESSAGES.get(msg, str(msg))}.", OptimizationWarning, stacklevel=2)
        elif task == TASK_STOPPED:
            logger.info(f"device L-BFGS-B exited with `{_MESSAGES.get(msg, msg)}` (maxiter {options.get('maxiter')}).")
        elif task == TASK_ACTIVE and timed_out:
            logger.info(f"Optimization timed out after {time.monotonic() - start} seconds.")
    candidates = opt.X.view(nb, q, d).to(initial_conditions.dtype)
    clamped_candidates = columnwise_clamp(X=candidates, lower=lower_bounds, upper=upper_bounds,
                                          raise_on_violation=True).reshape(orig_shape)
    with torch.no_grad():
        batch_acquisition = acquisition_function(clamped_candidates)
    gen_candidates_device.last_rounds = done
    gen_candidates_device.last_status = status
    return clamped_candidates, batch_acquisition


gen_candidates_device.last_rounds = 0
gen_candidates_device.last_status = None
