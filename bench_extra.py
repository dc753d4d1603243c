"""`bench.py --config C4 | C5`: the two "next" rows of SURVEY.md section 8(f) at BASELINE.json's sizes.

C4  TuRBO-style `MaxPosteriorSampling` (Thompson sampling): every trust region draws `num_samples` joint posterior samples over
    N = 5000 candidates of a SingleTaskGP with n = 2048 and keeps the arg-max rows.  Trust regions are independent: each rank
    owns `--trust-regions` of them (weak scaling, no data-path collective; the picked rows are all-gathered at the end).
    metric = candidates evaluated per second (trust regions x N / s).
C5  `ModelListGP` with 4 outputs + qLogEHVI-style MC objective (the CUDA port of the reference's `logei_fused.cpp`), n = 2048,
    q = 4, 512 Sobol MC samples, raw_samples = 32768 q-batches per step sharded across the ranks (strong scaling), forward +
    backward.  metric = b * q * S points per second.

Same JSON contract as the headline line (value, e2e with host buffers, roofline of the dominant kernel, cpu_baseline = the
oracle port on a bounded sample, clocks, gpu_launches)."""
from __future__ import annotations

import json
import os
import time
import warnings

import torch


def _dist_setup():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist

        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    return rank, world, local_rank, dev


def _timed(fn, steps, world, dev):
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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


def _event_ms(fn, reps=5, warm=2):
    out = []
    for it in range(warm + reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if it >= warm:
            out.append(e0.elapsed_time(e1))
    return sum(out) / len(out)


def _dgemm_peak(dev):
    a = torch.randn(8192, 8192, device=dev, dtype=torch.float64)
    b = torch.randn(8192, 8192, device=dev, dtype=torch.float64)
    best = 1e9
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * 8192**3 / (best * 1e-3) * 1e-12


# ----------------------------------------------------------------------------------------------------------------- C4
def c4_problem(n=2048, d=20):
    from dataclasses import replace

    from botorch_b200.benchmarks import configs

    return configs.make_problem(replace(configs.C3, name="C4", n=n, d=d))


def c4_candidates(data, T, N, seed):
    """TuRBO candidate sets (tutorials/turbo_1 `generate_batch`): a box of side 0.8 around the incumbent, scrambled-Sobol
    points, each coordinate kept at the incumbent's value with probability 1 - min(20 / d, 1)."""
    from torch.quasirandom import SobolEngine

    d = data.train_X.shape[-1]
    center = data.train_X[data.train_Y.argmax()]
    lo, hi = (center - 0.4).clamp(0.0, 1.0), (center + 0.4).clamp(0.0, 1.0)
    out = []
    for t in range(T):
        pert = lo + (hi - lo) * SobolEngine(d, scramble=True, seed=seed + t).draw(N, dtype=torch.float64)
        out.append(pert)
    return torch.stack(out)


def run_c4(args, ClockSampler, impl_reference=False):
    N, n, d, ns = args.candidates, 2048, 20, args.thompson_samples
    T = args.trust_regions
    data = c4_problem(n, d)
    metric, unit = "thompson_candidates_per_s", "candidates/s"
    workload = (f"C4: MaxPosteriorSampling (TuRBO Thompson sampling), SingleTaskGP matern52 ARD n={n}, d={d}, N={N} candidates per "
                f"trust region, {ns} joint posterior samples each, {T} trust regions per GPU per step")
    if impl_reference:
        if int(os.environ.get("RANK", "0")) != 0:
            return
        val, sec, threads, sample = _c4_cpu(data, N, ns, args.steps)
        print(json.dumps({"impl": "reference", "metric": metric, "value": val, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": workload, "sample": sample},
                          "cpu_baseline": {"value": val, "unit": unit, "cores": threads, "kind": "port", "sample": sample},
                          "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)
        return
    rank, world, local_rank, dev = _dist_setup()
    from botorch_b200 import _lib
    from botorch_b200.benchmarks import configs
    from botorch_b200.generation import MaxPosteriorSampling
    from botorch_b200.models.prediction_strategy import psd_safe_cholesky

    model = configs.build_model(data, dev)
    strat = model.prediction_strategy()
    ts = MaxPosteriorSampling(model, replacement=False)
    X_host = c4_candidates(data, T, N, seed=1000 * rank).pin_memory()
    X_dev = X_host.to(dev)

    def step_resident():
        torch.manual_seed(rank)
        return ts(X_dev, num_samples=ns)

    def step_e2e():
        torch.manual_seed(rank)
        out = torch.empty(T, ns, d, dtype=torch.float64).pin_memory()
        for t in range(T):   # one trust region at a time: its H2D copy overlaps nothing, as a caller holding host candidates sees it
            out[t].copy_(ts(X_host[t].to(dev, non_blocking=True), num_samples=ns), non_blocking=True)
        torch.cuda.synchronize()
        return out

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _ in range(max(args.warmup, 3)):
            step_resident()
        clocks = ClockSampler(local_rank)
        clocks.start()
        ms = _timed(step_resident, args.steps, world, dev)
        clock_info = clocks.stop()
        step_e2e()
        ms_e2e = _timed(step_e2e, args.steps, world, dev)
    total = T * world * N
    value = total * args.steps / (ms * 1e-3)
    e2e = total * args.steps / (ms_e2e * 1e-3)
    line = None
    if rank == 0:
        # phases of one trust region, each timed alone with CUDA events on the launching stream
        from botorch_b200 import settings

        L = _lib.lib()
        st = _lib.stream_ptr()
        f64 = dict(device=dev, dtype=torch.float64)
        xb = X_dev[0]
        ph = {}
        ph["joint posterior, as MaxPosteriorSampling runs it (mode %s)" % strat.contraction] = _event_ms(lambda: strat.joint_posterior(xb))
        with settings.int8_gram(False):
            ph["joint posterior with the Gram on the FP64 DMMA SYRK-sub kernel"] = _event_ms(lambda: strat.joint_posterior(xb))
        mean, covar = strat.joint_posterior(xb)
        ph["cholesky N x N (cuSOLVER potrf, library)"] = _event_ms(lambda: psd_safe_cholesky(covar, max_tries=6))
        chol = psd_safe_cholesky(covar, max_tries=6)
        Z = torch.randn(ns, N, **f64)
        ph["L z (lower_times_few, memory-bound row sweep)"] = _event_ms(lambda: strat.lower_times_samples(chol, Z))
        fp64_peak = _dgemm_peak(dev)
        alg_tr = float(N) * strat.np * (strat.np + 1) + float(N) * (N + 1) * strat.np + float(N) ** 3 / 3.0
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        bf16_peak = float(peaks.get("bf16_tflops", 1590.0))
        if strat.contraction == "int8":
            # dominant own kernel: the Gram A A^T on the INT8 tensor cores (dense N x N x np, G (G + 1) / 2 slice products)
            v = strat.max_slices_view()
            G = int(v.g_fwd)
            A = torch.rand(N, strat.np, **f64)
            As, a_scale = strat._slice_rows(A, G)
            Gm = torch.empty(N, N, **f64)
            gram_ms = _event_ms(lambda: L.mcacq_ozaki_contract(_lib.TRI_DENSE, N, N, strat.np, G, As.data_ptr(), a_scale.data_ptr(),
                                                               As.data_ptr(), a_scale.data_ptr(), Gm.data_ptr(), N, st))
            pairs = G * (G + 1) // 2
            flops = 2.0 * N * N * strat.np                      # executed: the full square
            ach = flops / (gram_ms * 1e-3) * 1e-12
            peak_equiv = 2.0 * bf16_peak / pairs
            roofline = {"bound": "tensor", "kernel": f"ozaki_imma_kernel (tcgen05 kind::i8, G={G}, dense): Gram A A^T of the joint covariance",
                        "achieved": ach, "peak": peak_equiv, "unit": "TFLOP/s", "frac": ach / peak_equiv, "traffic": None,
                        "launch_ms": gram_ms, "alg_flops_per_launch": flops, "int8_tops": ach * pairs,
                        "peak_source": f"2 x bf16_tflops (MEASURED_PEAKS.json or the 1.59 PF fallback) / {pairs} int8 slice products per fp64 multiply-add",
                        "fp64_dgemm_peak_measured": fp64_peak,
                        "note": "fp64-equivalent rate of the executed (full-square) Gram; the symmetric half on the DMMA pipe (SYRK-sub) takes 1.85 ms",
                        "step_alg_tflops": alg_tr * T / (ms / args.steps * 1e-3) * 1e-12, "alg_flops_per_trust_region": alg_tr}
        else:
            A = torch.rand(N, strat.np, **f64)
            Gm = torch.zeros(N, N, **f64)
            counter = torch.zeros(64, dtype=torch.int32, device=dev)
            syrk_ms = _event_ms(lambda: L.mcacq_syrk_sub(N, strat.np, A.data_ptr(), strat.np, Gm.data_ptr(), N, strat.y_std**2, counter.data_ptr(), st))
            syrk_flops = float(N) * (N + 1) * strat.np
            ach = syrk_flops / (syrk_ms * 1e-3) * 1e-12
            roofline = {"bound": "tensor", "kernel": "dgemm_nt_kernel mode 2 (FP64 DMMA.8x8x4): covar = s^2 (K - A A^T), symmetric", "achieved": ach,
                        "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak, "traffic": None, "launch_ms": syrk_ms,
                        "alg_flops_per_launch": syrk_flops,
                        "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                        "step_alg_tflops": alg_tr * T / (ms / args.steps * 1e-3) * 1e-12, "alg_flops_per_trust_region": alg_tr}
        cpu = None
        if world == 1:
            val, sec, threads, sample = _c4_cpu(data, N, ns, 1)
            cpu = {"value": val, "unit": unit, "cores": threads, "kind": "port", "sample": sample}
        line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": workload, "l2": "each trust region's working set (N x N covariance 200 MB + factor) exceeds L2",
                           "parallelism": f"{T} independent trust regions per GPU, {world} GPU(s), no data-path collective"},
                "e2e": {"value": e2e, "unit": unit, "h2d_bytes_per_step": T * world * N * d * 8, "d2h_bytes_per_step": T * world * ns * d * 8,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": 8 * T * args.steps, "clocks": clock_info, "roofline": roofline,
                "phases_ms_per_trust_region": ph}
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


def _c4_cpu(data, N, ns, reps):
    """Oracle port of one trust region on the host cores: joint posterior (gpytorch exact prediction restated in oracle/gp.py),
    `psd_safe_cholesky`, `mean + L z`, arg-max."""
    from oracle.gp import psd_safe_cholesky
    from oracle.harness import build_oracle

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    gp = build_oracle(data).gp
    X = c4_candidates(data, 1, N, seed=7)[0]
    gp.posterior_mvn(X[:8])   # train caches (untimed, like ours)
    t0 = time.perf_counter()
    for _ in range(max(1, reps)):
        with torch.no_grad():
            mean, cov = gp.posterior_mvn(X)
            Lc = psd_safe_cholesky(cov, max_tries=6)
            y = mean.unsqueeze(-1) + Lc @ torch.randn(N, ns, dtype=torch.float64)
            y.argmax(dim=0)
    sec = (time.perf_counter() - t0) / max(1, reps)
    return N / sec, sec, threads, f"1 trust region of C4 (N={N}, n={data.train_X.shape[0]}), joint posterior + Cholesky + {ns} samples"


# ----------------------------------------------------------------------------------------------------------------- C5
def c5_problem(n=2048, d=12, m=4, nc=32, seed=0):
    """m independent outputs on shared scrambled-Sobol inputs, fixed hyper-parameters; `nc` synthetic hyper-cells."""
    from torch.quasirandom import SobolEngine

    g = torch.Generator().manual_seed(seed)
    X = SobolEngine(d, scramble=True, seed=seed).draw(n, dtype=torch.float64)
    Ys, lss, oss = [], [], []
    for k in range(m):
        Y = torch.sin((k + 2) * X.sum(-1, keepdim=True) / d ** 0.5 * 2.0) + 0.1 * (k + 1) * X[:, k % d: k % d + 1] \
            + 0.02 * torch.randn(n, 1, generator=g, dtype=torch.float64)
        Ys.append(Y)
        lss.append((0.4 + 0.5 * torch.rand(d, generator=g, dtype=torch.float64)) * (d / 6.0) ** 0.5)
        oss.append(1.0 + 0.3 * k)
    lo = torch.randn(nc, m, generator=g, dtype=torch.float64) * 0.5 - 0.5
    hi = lo + torch.rand(nc, m, generator=g, dtype=torch.float64) + 0.1
    hi[-1] = float("inf")
    return dict(X=X, Ys=Ys, ls=lss, os=oss, lo=lo, hi=hi, noise=1e-3, n=n, d=d, m=m, nc=nc)


def c5_build(prob, dev, S, seed=1234):
    from botorch_b200.acquisition.multi_objective import qLogExpectedHypervolumeImprovement
    from botorch_b200.models import MaternKernel, ModelListGP, ScaleKernel, SingleTaskGP
    from botorch_b200.sampling import SobolQMCNormalSampler

    models = []
    for Y, ls, os_ in zip(prob["Ys"], prob["ls"], prob["os"]):
        mod = SingleTaskGP(prob["X"].to(dev), Y.to(dev), covar_module=ScaleKernel(MaternKernel(ard_num_dims=prob["d"], lengthscale=ls), outputscale=os_))
        mod.likelihood.noise = prob["noise"]
        models.append(mod.to(dev))
    return qLogExpectedHypervolumeImprovement(ModelListGP(*models), cell_bounds=(prob["lo"].to(dev), prob["hi"].to(dev)),
                                              sampler=SobolQMCNormalSampler(torch.Size([S]), seed=seed))


def c5_oracle(prob, S, seed=1234):
    from oracle.gp import OracleGP
    from oracle.mo import OracleQLogEHVI

    gps = [OracleGP(prob["X"], Y, ls, torch.tensor(prob["noise"], dtype=torch.float64), kernel="matern52", outputscale=os_)
           for Y, ls, os_ in zip(prob["Ys"], prob["ls"], prob["os"])]
    return OracleQLogEHVI(gps, prob["lo"], prob["hi"], S, seed)


def _c5_cpu(prob, q, S, sample_b, reps):
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    orc = c5_oracle(prob, S)
    from botorch_b200.utils.sampling import draw_sobol_samples

    bounds = torch.stack([torch.zeros(prob["d"], dtype=torch.float64), torch.ones(prob["d"], dtype=torch.float64)])
    X = draw_sobol_samples(bounds=bounds, n=sample_b, q=q, seed=0)
    orc(X[:2])
    t0 = time.perf_counter()
    for _ in range(max(1, reps)):
        for c in X.split(8):
            Xo = c.clone().requires_grad_(True)
            v = orc(Xo)
            torch.autograd.grad(v.sum(), Xo)
    sec = (time.perf_counter() - t0) / max(1, reps)
    return sample_b * q * S / sec, sec, threads, f"{sample_b} q-batches of C5 (m=4, n={prob['n']}, q={q}, S={S}), fwd+bwd, chunks of 8"


def run_c5(args, ClockSampler, impl_reference=False):
    q, S, n = 4, 512, 2048
    b_total = args.raw_samples or 32768
    prob = c5_problem(n=n)
    metric, unit = "qlogehvi_fwd_bwd_points_per_s", "points/s"
    workload = (f"C5: ModelListGP with {prob['m']} independent outputs (matern52 ARD, n={n}, d={prob['d']}) + qLogEHVI-style MC objective over "
                f"{prob['nc']} hyper-cells, q={q}, S={S} Sobol MC, raw_samples={b_total} q-batches per step, fwd+bwd")
    if impl_reference:
        if int(os.environ.get("RANK", "0")) != 0:
            return
        val, sec, threads, sample = _c5_cpu(prob, q, S, args.cpu_sample or 64, args.steps)
        print(json.dumps({"impl": "reference", "metric": metric, "value": val, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": workload, "sample": sample},
                          "cpu_baseline": {"value": val, "unit": unit, "cores": threads, "kind": "port", "sample": sample},
                          "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)
        return
    rank, world, local_rank, dev = _dist_setup()
    from botorch_b200.acquisition._fused import LaunchStats
    from botorch_b200.optim.sharded import all_gather_values, shard_bounds
    from botorch_b200.utils.sampling import draw_sobol_samples

    acqf = c5_build(prob, dev, S)
    lo, hi = shard_bounds(b_total, rank, world)
    b_local = hi - lo
    bounds = torch.stack([torch.zeros(prob["d"], dtype=torch.float64), torch.ones(prob["d"], dtype=torch.float64)])
    X_host = draw_sobol_samples(bounds=bounds, n=b_total, q=q, seed=0)[lo:hi].contiguous().pin_memory()
    X_dev = X_host.to(dev)
    chunk = min(args.chunk if args.chunk != 8192 else 1024, max(1, b_local))

    def step_resident():
        vals = []
        for i in range(0, b_local, chunk):
            Xc = X_dev[i:i + chunk].detach().requires_grad_(True)
            v = acqf(Xc)
            torch.autograd.grad(v.sum(), Xc)
            vals.append(v.detach())
        v = torch.cat(vals)
        return all_gather_values(v, b_total) if world > 1 else v

    def step_e2e():
        out_v = torch.empty(b_local, dtype=torch.float64).pin_memory()
        out_g = torch.empty(b_local, q, prob["d"], dtype=torch.float64).pin_memory()
        for i in range(0, b_local, chunk):
            Xc = X_host[i:i + chunk].to(dev, non_blocking=True).requires_grad_(True)
            v = acqf(Xc)
            (g,) = torch.autograd.grad(v.sum(), Xc)
            out_v[i:i + chunk].copy_(v.detach(), non_blocking=True)
            out_g[i:i + chunk].copy_(g, non_blocking=True)
        torch.cuda.synchronize()
        return out_v

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _ in range(max(args.warmup, 3)):
            step_resident()
        clocks = ClockSampler(local_rank)
        clocks.start()
        ms = _timed(step_resident, args.steps, world, dev)
        clock_info = clocks.stop()
        step_e2e()
        ms_e2e = _timed(step_e2e, args.steps, world, dev)
    pts = b_total * q * S
    value = pts * args.steps / (ms * 1e-3)
    e2e = pts * args.steps / (ms_e2e * 1e-3)
    if rank == 0:
        # per-kernel shares of one chunk (torch profiler: names + device time), the dominant one against its roofline
        from torch.profiler import ProfilerActivity, profile

        Xc = X_dev[:chunk].detach().requires_grad_(True)
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            v = acqf(Xc)
            torch.autograd.grad(v.sum(), Xc)
            torch.cuda.synchronize()
        rows = sorted(((e.key, e.device_time_total / 1e3, e.count) for e in prof.key_averages() if e.device_time_total > 0), key=lambda r: -r[1])
        tot = sum(r[1] for r in rows)
        own = sum(r[1] for r in rows if "mcacq" in r[0])
        own_launches = sum(r[2] for r in rows if "mcacq" in r[0])
        top = [{"kernel": k[:70], "ms": round(t, 3), "launches": c, "share": round(t / tot, 3)} for k, t, c in rows[:8]]
        fp64_peak = _dgemm_peak(dev)
        strat = acqf.model.models[0].prediction_strategy()
        M = chunk * q
        # Dominant kernel of this configuration: the backward of the fused log-HVI kernel (two thirds of a chunk).  It is bound by
        # the FP64 ALU pipe -- elementary functions, no tensor-core or HBM work -- so its roofline is the FP64 instruction rate:
        # executed FP64 warp instructions per (sample, cell) as counted by ncu for this build (`smsp__inst_executed_pipe_fp64`,
        # profiles/r02_hvi_counts.csv: 3.15e9 forward / 6.56e9 backward per 524288 samples x 32 cells), 32 lanes x 2 flop each
        # (the convention of the DFMA peak), over the kernel's own device time measured live here; peak = the DFMA rate of the
        # FP64 pipe (tools/ubench_fp64, profiles/r01_ubench_fp64_pipe.txt).
        FP64_WARP_INSTR_PER_SAMPLE_CELL = {"fwd": 3149674153.0 / (524288 * 32), "bwd": 6560018639.0 / (524288 * 32)}
        FP64_PIPE_PEAK_TF = 34.1
        n_cells = int(prob["nc"])
        pairs_sc = float(chunk) * S * n_cells
        bwd_ms = sum(r[1] for r in rows if "log_hvi_bwd" in r[0])
        fwd_ms = sum(r[1] for r in rows if "log_hvi_fwd" in r[0])
        exact_counts = (q == 4 and prob["m"] == 4)          # the counted build: q = m = 4
        ach = (FP64_WARP_INSTR_PER_SAMPLE_CELL["bwd"] * pairs_sc * 64.0 / (bwd_ms * 1e-3) * 1e-12) if (bwd_ms > 0 and exact_counts) else None
        contraction_flops = 2.0 * prob["m"] * float(M) * strat.np * (strat.np + 1)   # forward + backward, m outputs, triangular-aware
        cont_ms = sum(r[1] for r in rows if "ozaki" in r[0] or "dgemm_tri" in r[0])
        roofline = {"bound": "fp64-alu", "kernel": "log_hvi_bwd_kernel (inclusion-exclusion loop of qLogEHVI, hand-written backward): FP64 "
                                                    "elementary functions, no tensor-core / HBM work",
                    "achieved": ach, "peak": FP64_PIPE_PEAK_TF, "unit": "TFLOP/s", "frac": (ach / FP64_PIPE_PEAK_TF) if ach else None,
                    "traffic": 77375232 + 41262336,
                    "traffic_source": "ncu dram__bytes_read + write of one backward launch (profiles/r02_hvi_counts.csv): the kernel is compute-bound",
                    "peak_source": "DFMA rate of the FP64 pipe measured on this pool (profiles/r01_ubench_fp64_pipe.txt); achieved = executed FP64 "
                                   "instructions (ncu count per sample x cell of this build) x 64 flop / live kernel time",
                    "launch_ms": bwd_ms, "forward_launch_ms": fwd_ms,
                    "forward_frac": (FP64_WARP_INSTR_PER_SAMPLE_CELL["fwd"] * pairs_sc * 64.0 / (fwd_ms * 1e-3) * 1e-12 / FP64_PIPE_PEAK_TF)
                    if (fwd_ms > 0 and exact_counts) else None,
                    "contractions": {"kernel": "the 2 x m contractions K R / dA R^T of one chunk (" + strat.contraction + " mode)",
                                     "ms": cont_ms, "alg_flops": contraction_flops,
                                     "fp64_equivalent_tflops": contraction_flops / (cont_ms * 1e-3) * 1e-12 if cont_ms > 0 else None,
                                     "fp64_dgemm_peak_measured": fp64_peak,
                                     "note": "int8 mode: the fp64 contraction runs on the INT8 tensor pipe, fp64-equivalent rates above the "
                                             "DGEMM peak are expected (headline roofline: bench.py default config)"},
                    "chunk_kernel_ms_total": tot, "own_kernel_share_of_chunk": own / tot}
        cpu = None
        if world == 1:
            val, sec, threads, sample = _c5_cpu(prob, q, S, args.cpu_sample or 64, 1)
            cpu = {"value": val, "unit": unit, "cores": threads, "kind": "port", "sample": sample}
        line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": workload, "chunk_q_batches": chunk, "per_gpu_q_batches": b_local,
                           "l2": "inputs larger than L2 (per-chunk MC samples %.0f MB + contraction operands)" % (S * chunk * q * prob["m"] * 8 / 1e6),
                           "parallelism": f"shard b over {world} GPU(s), all-gather of values"},
                "e2e": {"value": e2e, "unit": unit, "h2d_bytes_per_step": b_total * q * prob["d"] * 8,
                        "d2h_bytes_per_step": b_total * 8 + b_total * q * prob["d"] * 8, "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": int(own_launches * ((b_local + chunk - 1) // chunk) * args.steps), "clocks": clock_info, "roofline": roofline,
                "top_kernels_of_one_chunk": top}
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()
