#!/usr/bin/env python
"""Benchmark of the hot path: qLogNEI forward+backward evaluations per second (b*q*MC points/s).

Contract: `python bench.py --gpus N --steps K --warmup W [--impl reference]`, one JSON line on stdout (rank 0).
A "step" is one forward+backward pass over the whole raw-sample sweep of the workload (b = raw_samples
q-batches, sharded across the N ranks: strong scaling), evaluated in chunks like `init_batch_limit`.
See DESIGN.md section "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import warnings

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "qlognei_fwd_bwd_points_per_s"
UNIT = "points/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=["C1", "C2", "C3", "C4", "C5"],
                    help="C1-C3: the qLogEI / qLogNEI hot path (C3 = the headline); C4: MaxPosteriorSampling (Thompson); "
                         "C5: ModelListGP + qLogEHVI-style objective (bench_extra.py)")
    ap.add_argument("--trust-regions", type=int, default=8, help="C4: trust regions per GPU per step")
    ap.add_argument("--candidates", type=int, default=5000, help="C4: candidates per trust region")
    ap.add_argument("--thompson-samples", type=int, default=4, help="C4: joint posterior samples (batch size) per trust region")
    ap.add_argument("--raw-samples", type=int, default=None, help="override b (total q-batches per step)")
    ap.add_argument("--chunk", type=int, default=8192, help="q-batches per fused call (init_batch_limit analogue)")
    ap.add_argument("--cpu-sample", type=int, default=None, help="q-batches in the CPU baseline sample")
    ap.add_argument("--contraction", default="int8", choices=["int8", "dmma"],
                    help="headline contraction mode: int8 = Ozaki split on the INT8 tensor cores (tcgen05), "
                         "dmma = FP64 DMMA kernel; the other mode is measured too and reported under 'alt_mode'")
    return ap.parse_args()


class ClockSampler:
    """Samples SM clocks / throttle reasons while the timed region runs (B200_PROFILING.md clocks line).

    NVML is polled from a thread every 5 ms (the timed region of a multi-GPU run is ~100 ms, too short for
    `nvidia-smi -lms`); `nvidia-smi --query-gpu` is the fallback when pynvml is missing."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None
        self.sm, self.mask, self.max_mhz, self.handle, self.nvml = [], 0, None, None, None
        self._stop = threading.Event()
        self.thread = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(self.index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        while True:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                try:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            except Exception:
                pass
            if self._stop.wait(0.005):
                return

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.thread.join(timeout=1.0)
            reasons = sorted(name for bit, name in self.BITS.items() if self.mask & bit)
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": reasons, "samples": len(self.sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower() == "active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": "nvidia-smi"}


REFERENCE_KIND = "port"   # becomes "botorch" when a real BoTorch + gpytorch install is importable (never in this image)


def _real_botorch_acqf(data, spec):
    """The UNMODIFIED reference on the same synthetic problem, if `botorch` + `gpytorch` import (site-packages or
    `baseline/_ref`); None otherwise.  (SURVEY.md fact 1 / VERDICT r01 item 1d: no wheels and no network here, so this
    branch has never run in this image; `tests/test_real_botorch_probe.py` is its correctness twin.)"""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(ref) and ref not in sys.path:
        sys.path.append(ref)
    try:
        import gpytorch  # noqa: F401
        from botorch.acquisition.logei import qLogExpectedImprovement, qLogNoisyExpectedImprovement
        from botorch.models import SingleTaskGP
        from botorch.models.transforms.outcome import Standardize
        from botorch.sampling.normal import SobolQMCNormalSampler
        from gpytorch.kernels import MaternKernel, RBFKernel, ScaleKernel
    except Exception:  # noqa: BLE001 -- not installed / not importable: the oracle port is timed instead
        return None
    base = (RBFKernel if spec.kernel == "rbf" else MaternKernel)(ard_num_dims=spec.d)
    base.lengthscale = data.lengthscale
    covar = base
    if spec.outputscale is not None:
        covar = ScaleKernel(base)
        covar.outputscale = spec.outputscale
    model = SingleTaskGP(data.train_X, data.train_Y, covar_module=covar, outcome_transform=Standardize(m=1))
    model.likelihood.noise = data.noise
    model.mean_module.constant = 0.0
    model.eval()
    sampler = SobolQMCNormalSampler(sample_shape=torch.Size([spec.S]), seed=1234)
    if spec.acqf == "qLogEI":
        return qLogExpectedImprovement(model, best_f=data.best_f, sampler=sampler)
    return qLogNoisyExpectedImprovement(model, X_baseline=data.X_baseline, sampler=sampler, prune_baseline=False)


def cpu_reference(args, data, spec, steps, warmup, sample_b):
    """The reference's CPU path on a bounded sample, all host threads: real BoTorch when importable (kind "botorch"), else the
    oracle port (pure-torch restatement, kind "port")."""
    global REFERENCE_KIND
    from oracle.harness import build_oracle, time_cpu_fwd_bwd
    from botorch_b200.benchmarks import configs

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    real = _real_botorch_acqf(data, spec)
    REFERENCE_KIND = "botorch" if real is not None else "port"
    orc = real if real is not None else build_oracle(data)
    X = configs.eval_points(data, sample_b)
    chunk = min(sample_b, 64 if spec.n >= 4096 else 256)
    from oracle.acquisition import value_and_grad
    value_and_grad(orc, X[:min(8, sample_b)])  # builds the train caches (one-off, untimed like ours)
    sec, threads = time_cpu_fwd_bwd(orc, X, chunk=chunk, warmup=1 if warmup else 0, reps=max(1, steps))
    pts = sample_b * spec.q * spec.S
    return pts / sec, sec, threads, f"{sample_b} q-batches of {spec.name} (n={spec.n}, q={spec.q}, S={spec.S}), fwd+bwd, chunks of {chunk}"


def main():
    args = parse()
    if args.config in ("C4", "C5"):
        import bench_extra

        if args.impl != "reference" and not torch.cuda.is_available():
            raise SystemExit("bench.py: CUDA is required (botorch_b200 has no CPU path)")
        (bench_extra.run_c4 if args.config == "C4" else bench_extra.run_c5)(args, ClockSampler, impl_reference=args.impl == "reference")
        return
    from botorch_b200.benchmarks import configs

    spec = configs.CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    b_total = args.raw_samples or spec.raw_samples
    workload = (f"{spec.name}: {spec.acqf} {spec.kernel} ARD, n={spec.n}, d={spec.d}, q={spec.q}, S={spec.S} Sobol MC, "
                f"r={spec.r} baseline pts, raw_samples={b_total} q-batches per step, fwd+bwd")
    data = configs.make_problem(spec)

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return
        sample_b = args.cpu_sample or (2048 if spec.n >= 4096 else 8192)   # ~8 s of host work per step on 16 cores
        val, sec, threads, sample = cpu_reference(args, data, spec, args.steps, args.warmup, sample_b)
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "sample": sample},
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": REFERENCE_KIND, "sample": sample},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return

    # ------------------------------------------------------------------ our arm
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: CUDA is required (botorch_b200 has no CPU path)")
    import torch.distributed as dist
    from botorch_b200 import _lib
    from botorch_b200.acquisition._fused import LaunchStats, RerouteStats
    from botorch_b200.optim.sharded import all_gather_values, shard_bounds

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner on STDOUT when NCCL_DEBUG=VERSION; the contract is ONE JSON line on stdout
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")   # (the banner is printed at every level >= VERSION)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    lo, hi = shard_bounds(b_total, rank, world)
    b_local = hi - lo

    from botorch_b200 import settings

    X_host = configs.eval_points(data, b_total)[lo:hi].contiguous().pin_memory()
    X_dev = X_host.to(dev)
    chunk = min(args.chunk, max(1, b_local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    pts_local_scale = b_total * spec.q * spec.S  # whole-job points per step (every rank sweeps its shard concurrently)

    def measure(mode: str, with_clocks: bool):
        """Build the model in the given contraction mode and time K resident steps and K host-buffer (e2e) steps."""
        with settings.contraction(mode):
            model = configs.build_model(data, dev)
            acqf = configs.build_acqf(data, model)
            model.prediction_strategy()

            def step_resident():
                vals = []
                for i in range(0, b_local, chunk):
                    Xc = X_dev[i:i + chunk].detach().requires_grad_(True)
                    v = acqf(Xc)
                    (g,) = torch.autograd.grad(v.sum(), Xc)
                    vals.append(v.detach())
                v = torch.cat(vals)
                return all_gather_values(v, b_total) if world > 1 else v

            # pinned result buffers are allocated once (as a caller sweeping many batches would); every step copies its inputs
            # host -> device and its values / gradients device -> host inside the timed region
            out_v = torch.empty(b_local, dtype=torch.float64).pin_memory()
            out_g = torch.empty(b_local, spec.q, spec.d, dtype=torch.float64).pin_memory()

            def step_e2e():
                for i in range(0, b_local, chunk):
                    Xc = X_host[i:i + chunk].to(dev, non_blocking=True).requires_grad_(True)
                    v = acqf(Xc)
                    (g,) = torch.autograd.grad(v.sum(), Xc)
                    out_v[i:i + chunk].copy_(v.detach(), non_blocking=True)
                    out_g[i:i + chunk].copy_(g, non_blocking=True)
                torch.cuda.synchronize()
                return out_v, out_g

            for _ in range(max(args.warmup, 3)):
                step_resident()
            LaunchStats.launches = 0
            RerouteStats.q_batches = RerouteStats.calls = 0
            clocks = ClockSampler(local_rank) if with_clocks else None
            if clocks:
                clocks.start()
            ms_total = timed(step_resident, args.steps)
            clock_info = clocks.stop() if clocks else None
            launches = LaunchStats.launches
            rerouted = RerouteStats.q_batches
            for _ in range(2):
                step_e2e()
            ms_e2e = timed(step_e2e, args.steps)
            phases = None
            if with_clocks:
                # the other two shapes SURVEY.md section 8(d) asks for: the raw-sample sweep without gradient, and one
                # L-BFGS round (b = num_restarts q-batches, forward + backward, a latency-bound call)
                def step_fwd():
                    with torch.no_grad():
                        for i in range(0, b_local, chunk):
                            acqf(X_dev[i:i + chunk])

                step_fwd()
                ms_fwd = timed(step_fwd, args.steps)
                nr = max(1, min(spec.num_restarts, b_local))

                def step_round():
                    # as gen_candidates_scipy evaluates a round (generation/gen.py): the int8 mode's most accurate slice counts
                    with settings.int8_max_slices(True):
                        Xc = X_dev[:nr].detach().requires_grad_(True)
                        v = acqf(Xc)
                        torch.autograd.grad(v.sum(), Xc)

                for _ in range(3):
                    step_round()
                ms_round = timed(step_round, 20) / 20
                opt = None
                if world == 1:
                    # the call a user makes: optimize_acqf over the unit box (raw-sample sweep, initial-condition selection,
                    # batched L-BFGS-B with maxiter 50, final arg-max), wall clock of the second call
                    import time as _time

                    from botorch_b200.optim import optimize_acqf

                    ob = torch.stack([torch.zeros(spec.d), torch.ones(spec.d)]).to(dev, torch.float64)
                    okw = dict(bounds=ob, q=spec.q, num_restarts=spec.num_restarts, raw_samples=spec.raw_samples,
                               options={"maxiter": 50, "seed": 0})
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        torch.manual_seed(0)   # the Boltzmann draw of the initial conditions uses the global RNG: same restarts,
                        optimize_acqf(acqf, **okw)   # hence the same number of rounds, in every call and every run
                        torch.cuda.synchronize()
                        def _timed_opt():
                            # wall clock of one call; the smaller of two (a single shot occasionally carries a ~0.2 s allocator
                            # / graph re-capture hiccup after the memory-heavy phases before it); both are reported
                            walls, val = [], None
                            for _ in range(2):
                                torch.manual_seed(0)
                                t0 = _time.perf_counter()
                                _, val = optimize_acqf(acqf, **okw)
                                torch.cuda.synchronize()
                                walls.append((_time.perf_counter() - t0) * 1e3)
                            return min(walls), walls, val

                        wall, walls, oval = _timed_opt()
                        opt = {"wall_ms": wall, "wall_ms_calls": walls, "q": spec.q, "num_restarts": spec.num_restarts,
                               "raw_samples": spec.raw_samples, "maxiter": 50, "acq_value": float(oval),
                               "optimizer": "scipy (default): scipy's setulb stepped on the host, one fused fwd+bwd per round"}
                        # the same call with the device-resident L-BFGS-B (settings.optimizer('device'): CUDA-graph rounds)
                        with settings.optimizer("device"):
                            torch.manual_seed(0)
                            optimize_acqf(acqf, **okw)
                            torch.cuda.synchronize()
                            wall_d, walls_d, oval_d = _timed_opt()
                            opt["device_optimizer"] = {"wall_ms": wall_d, "wall_ms_calls": walls_d, "acq_value": float(oval_d)}
                phases = {"optimize_acqf": opt,
                          "sweep_forward_only_points_per_s": pts_local_scale * args.steps / (ms_fwd * 1e-3),
                          "sweep_forward_only_ms_per_step": ms_fwd / args.steps,
                          "lbfgs_round": {"q_batches": nr, "ms_per_fwd_bwd_call": ms_round,
                                          "note": "per rank, not sharded: one optimiser round over num_restarts q-batches"}}
        return {"model": model, "ms_total": ms_total, "ms_e2e": ms_e2e, "launches": launches, "rerouted": rerouted, "clocks": clock_info,
                "phases": phases}

    other = "dmma" if args.contraction == "int8" else "int8"
    head = measure(args.contraction, with_clocks=True)
    alt = measure(other, with_clocks=False)
    model = head["model"]
    ms_total, ms_e2e, launches, clock_info = head["ms_total"], head["ms_e2e"], head["launches"], head["clocks"]
    pts_per_step = b_total * spec.q * spec.S
    value = pts_per_step * args.steps / (ms_total * 1e-3)
    e2e_value = pts_per_step * args.steps / (ms_e2e * 1e-3)

    # ------------------------------------------------------------------ roofline of the dominant kernel (rank 0)
    roofline, roofline_alt, cpu_base = None, None, None
    if rank == 0:
        strat = model.prediction_strategy()
        L = _lib.lib()
        M = chunk * spec.q
        f64 = dict(device=dev, dtype=torch.float64)
        st = _lib.stream_ptr()
        alg_flops = float(M) * strat.np * (strat.np + 1)  # 2 * M * np*(np+1)/2: only the triangle of R is contracted

        def time_launch(fn, reps):
            durs = []
            for it in range(3 + reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn(it)
                e1.record()
                torch.cuda.synchronize()
                if it >= 3:
                    durs.append(e0.elapsed_time(e1))
            return sum(durs) / len(durs)

        # FP64 peak: MEASURED_PEAKS.json carries HBM and bf16 only, so the FP64 tensor denominator is measured here with
        # cuBLAS DGEMM 8192^3 (best of 6), as SURVEY.md section 8d prescribes.
        a = torch.randn(8192, 8192, **f64)
        bmat = torch.randn(8192, 8192, **f64)
        best = 1e9
        for _ in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, bmat)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        fp64_peak_tf = 2.0 * 8192**3 / (best * 1e-3) * 1e-12
        del a, bmat
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        bf16_peak = float(peaks.get("bf16_tflops", 1590.0))
        peak_src = "MEASURED_PEAKS.json" if "bf16_tflops" in peaks else "fallback 1.59 PF (B200_PROFILING.md)"

        def dmma_roofline():
            A = torch.randn(M, strat.np, **f64)
            Cc = torch.empty(M, strat.np, **f64)
            counter = torch.zeros(64, dtype=torch.int32, device=dev)

            def fn(it):
                mode, Bm = (_lib.TRI_UPPER, strat.R) if it % 2 == 0 else (_lib.TRI_LOWER, strat.Rt)
                L.mcacq_dgemm_tri(mode, M, strat.np, A.data_ptr(), Bm.data_ptr(), Cc.data_ptr(), counter.data_ptr(), st)

            ms = time_launch(fn, 2 * args.steps)
            ach = alg_flops / (ms * 1e-3) * 1e-12
            return {"bound": "tensor", "kernel": "dgemm_tri_kernel (FP64 DMMA.8x8x4)", "achieved": ach, "peak": fp64_peak_tf,
                    "unit": "TFLOP/s", "frac": ach / fp64_peak_tf, "traffic": 3.16e9 * (M / 16384.0),
                    "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no FP64 entry; "
                                   "DMMA pipe microbenchmark 37.2 TF/s, profiles/r01_ubench_fp64_pipe.txt)",
                    "launch_ms": ms, "alg_flops_per_launch": alg_flops,
                    "traffic_source": "ncu dram__bytes_read+write at M=16384 scaled to this M (profiles/README.md)"}

        def int8_roofline():
            st8 = head["model"].prediction_strategy() if args.contraction == "int8" else alt["model"].prediction_strategy()
            G = int(st8.g_fwd) if st8.contraction == "int8" else 7   # the slice count the build-time probe picked for THIS model
            pairs = G * (G + 1) // 2
            A = torch.rand(M, strat.np, **f64)
            sl = torch.empty(G, M, strat.np, dtype=torch.int8, device=dev)
            rs = torch.empty(M, **f64)
            L.mcacq_slice_rows(A.data_ptr(), M, strat.np, strat.np, strat.np, G, 1, 1, sl.data_ptr(), rs.data_ptr(), st)
            Cc = torch.empty(M, strat.np, **f64)
            Bs, bs = strat._slice_rows(strat.Rt, G)

            def fn(it):
                L.mcacq_ozaki_contract(_lib.TRI_UPPER, M, strat.np, strat.np, G, sl.data_ptr(), rs.data_ptr(), Bs.data_ptr(),
                                       bs.data_ptr(), Cc.data_ptr(), strat.np, st)

            ms = time_launch(fn, 2 * args.steps)
            ach = alg_flops / (ms * 1e-3) * 1e-12
            # what the vendor library reaches for int8 x int8 -> int32 on this part (cuBLASLt via torch._int_mm, 8192^3,
            # best of 5): 2.5-2.9 POP/s measured, i.e. 1.55-1.8x the bf16 rate rather than the nominal 2x
            lib_int8 = None
            try:
                ai = torch.randint(-128, 127, (8192, 8192), dtype=torch.int8, device=dev)
                bi = torch.randint(-128, 127, (8192, 8192), dtype=torch.int8, device=dev)
                best_i = 1e9
                for _ in range(6):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    torch._int_mm(ai, bi)
                    e1.record()
                    torch.cuda.synchronize()
                    best_i = min(best_i, e0.elapsed_time(e1))
                lib_int8 = 2.0 * 8192**3 / (best_i * 1e-3) * 1e-12
                del ai, bi
            except Exception:  # noqa: BLE001 -- informational only
                lib_int8 = None
            # roofline of THIS algorithm on the INT8 tensor pipe: every fp64 multiply-add costs G(G+1)/2 int8 multiply-adds;
            # dense int8 peak of the part = 2 x the measured dense bf16 peak
            peak_equiv = 2.0 * bf16_peak / pairs
            return {"bound": "tensor", "kernel": f"ozaki_imma_kernel (tcgen05 kind::i8, G={G} forward launch)", "achieved": ach,
                    "slices_fwd_bwd": [int(st8.g_fwd), int(st8.g_bwd)], "mode_in_effect": st8.contraction,
                    "peak": peak_equiv, "unit": "TFLOP/s", "frac": ach / peak_equiv, "traffic": 8.93e9 * (M / 65536.0) * (G / 6.0),
                    "peak_source": f"2 x bf16_tflops ({peak_src}) / {pairs} int8 slice products per fp64 multiply-add",
                    "launch_ms": ms, "alg_flops_per_launch": alg_flops, "int8_tops": ach * pairs,
                    "library_int8_gemm_tops_measured": lib_int8,
                    "frac_of_library_int8_gemm": (ach * pairs / lib_int8) if lib_int8 else None,
                    "fp64_dgemm_peak_measured": fp64_peak_tf,
                    "traffic_source": "ncu --set full dram__bytes_read+write of the G=6 forward launch at M=65536 (6.80 + 2.13 GB, "
                                      "profiles/r02_ncu_full_c3_chunk_6_5_raw.csv; algorithmic: 1.7 GB slices + 2.15 GB output), "
                                      "scaled to this M and G"}

        roof = {"int8": int8_roofline, "dmma": dmma_roofline}
        roofline = roof[args.contraction]()
        roofline_alt = roof[other]()
        roofline["step_alg_tflops"] = (2.0 * spec.q * spec.n * spec.n * b_total / world) / (ms_total / args.steps * 1e-3) * 1e-12
        if world == 1:
            sample_b = args.cpu_sample or (2048 if spec.n >= 4096 else 8192)  # ~10 s of host work (bounded sample)
            val, sec, threads, sample = cpu_reference(args, data, spec, 1, 1, sample_b)
            cpu_base = {"value": val, "unit": UNIT, "cores": threads, "kind": REFERENCE_KIND, "sample": sample}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "chunk_q_batches": chunk, "per_gpu_q_batches": b_local,
                           "l2": "inputs larger than L2 (per-chunk working set %.1f GB)" % (2 * chunk * spec.q * model.prediction_strategy().np * 8 / 1e9),
                           "parallelism": f"shard b over {world} GPU(s), all-gather of values",
                           "contraction": (("int8 (library default): Ozaki split of the fp64 contraction onto the INT8 tensor cores "
                                            "(tcgen05); %d forward / %d backward signed 8-bit slices picked by the per-model probe; "
                                            "q-batches whose variance has collapsed below %.3g of the prior (at / next to training "
                                            "points) or whose conditional covariance is nearly singular are re-evaluated through the "
                                            "FP64 DMMA contraction: %d of %d q-batches in the timed steps"
                                            % (model.prediction_strategy().g_fwd, model.prediction_strategy().g_bwd,
                                               model.prediction_strategy().int8_var_ratio_limit or 0.0, head["rerouted"],
                                               b_total * args.steps))
                                           if model.prediction_strategy().contraction == "int8" else "dmma: FP64 DMMA tensor-core kernel")},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": b_total * spec.q * spec.d * 8,
                        "d2h_bytes_per_step": b_total * 8 + b_total * spec.q * spec.d * 8, "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches, "clocks": clock_info, "roofline": roofline, "phases": head["phases"],
                "alt_mode": {"contraction": other, "value": pts_per_step * args.steps / (alt["ms_total"] * 1e-3),
                             "ms_per_step": alt["ms_total"] / args.steps,
                             "e2e_value": pts_per_step * args.steps / (alt["ms_e2e"] * 1e-3), "roofline": roofline_alt}}
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
