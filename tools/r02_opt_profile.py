"""Developer tool: where the time of gen_candidates_device goes (CUDA events around graph replays / the step kernel)."""
import sys, time, warnings
import torch
sys.path.insert(0, ".")
from botorch_b200.benchmarks import configs
from botorch_b200.generation.device_gen import DeviceLBFGSB, _FusedRound
dev = torch.device("cuda:0"); warnings.simplefilter("ignore")
for name in sys.argv[1:] or ["C1", "C3"]:
    spec = configs.CONFIGS[name]; data = configs.make_problem(spec)
    model = configs.build_model(data, dev); acqf = configs.build_acqf(data, model)
    ics = configs.eval_points(data, spec.num_restarts, seed=5).to(dev)
    nb, q, d = ics.shape
    acqf(ics)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    opt = DeviceLBFGSB(ics.reshape(nb, q * d), torch.zeros(q * d, device=dev, dtype=torch.float64), torch.ones(q * d, device=dev, dtype=torch.float64), maxiter=50)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    rnd = _FusedRound(acqf, opt, q, d, use_graph=True)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{name}: init {1e3*(t1-t0):.2f} ms, warm-up + capture {1e3*(t2-t1):.2f} ms, graph={'yes' if rnd.graph is not None else 'NO'} launches/round {rnd.launches_per_round}")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(40): rnd.run()
    e1.record(); torch.cuda.synchronize()
    print(f"   graph replay: {e0.elapsed_time(e1)/40:.3f} ms/round (device time)")
    t0 = time.perf_counter()
    for _ in range(40): rnd.run()
    torch.cuda.synchronize()
    print(f"   graph replay: {(time.perf_counter()-t0)*1e3/40:.3f} ms/round (wall)")
    # step kernel alone on the current state (re-init to keep problems active)
    opt2 = DeviceLBFGSB(ics.reshape(nb, q * d), torch.zeros(q * d, device=dev, dtype=torch.float64), torch.ones(q * d, device=dev, dtype=torch.float64), maxiter=500)
    rnd2 = _FusedRound(acqf, opt2, q, d, use_graph=False)
    for _ in range(5): rnd2.run()
    times = []
    for _ in range(20):
        rnd2._launch.__self__  # noqa
        # forward+backward then timed step
        import ctypes as C
        from botorch_b200 import _lib
        L, st = _lib.lib(), _lib.stream_ptr()
        base = C.byref(rnd2.base.desc) if rnd2.base is not None else None
        L.mcacq_acq_forward(C.byref(rnd2.strat.desc), base, C.byref(rnd2.mc.desc), opt2.X.data_ptr(), nb, q, rnd2.acq.data_ptr(), rnd2.info.data_ptr(), rnd2.ws.data_ptr(), rnd2.ws.numel(), st)
        L.mcacq_acq_backward(C.byref(rnd2.strat.desc), base, C.byref(rnd2.mc.desc), opt2.X.data_ptr(), nb, q, rnd2.acq.data_ptr(), rnd2.ones.data_ptr(), rnd2.gX.data_ptr(), rnd2.ws.data_ptr(), rnd2.ws.numel(), st)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); opt2.step(rnd2.acq, rnd2.gX.view(nb, -1), sign=-1.0); b.record(); torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    print(f"   step kernel alone: median {sorted(times)[10]*1e3:.1f} us, max {max(times)*1e3:.1f} us, active {opt2.active()}")
