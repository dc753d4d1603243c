"""Developer experiment (GPU): how much of the SIMT work (covariance, blocks, sample/reduce: FP64 pipe) hides under the int8
tensor-core contraction of ANOTHER chunk when consecutive chunks run on two streams with preallocated workspaces and no host
synchronisation -- straight through the C ABI.  Prints ms per 8192-q-batch chunk (forward + backward) for 1, 2, 3 streams."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from botorch_b200 import _lib, settings
from botorch_b200.benchmarks import configs

settings.contraction.set(os.environ.get("MCACQ_CONTRACTION", "int8"))
dev = torch.device("cuda:0")
cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
b = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
spec = configs.CONFIGS[cfg]
data = configs.make_problem(spec); model = configs.build_model(data, dev); acqf = configs.build_acqf(data, model)
strat = model.prediction_strategy()
nchunks = 8
Xs = [configs.eval_points(data, b).to(dev).contiguous() for _ in range(nchunks)]
base = acqf._baseline_operands() if hasattr(acqf, "_baseline_operands") else None
mc = acqf._mc_operands(Xs[0])
L = _lib.lib()
q = spec.q
f64 = dict(device=dev, dtype=torch.float64)


def run(nstreams, fwd_only=False):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    wss = [strat.workspace(b, q, base.r if base is not None else 0) for _ in range(nstreams)]
    acqs = [torch.empty(b, **f64) for _ in range(nchunks)]
    infos = [torch.empty(b, dtype=torch.int32, device=dev) for _ in range(nchunks)]
    gXs = [torch.empty_like(Xs[0]) for _ in range(nchunks)]
    ones = torch.ones(b, **f64)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams: s.wait_stream(torch.cuda.current_stream())
    for i in range(nchunks):
        s = streams[i % nstreams]; ws = wss[i % nstreams]
        sp = C.c_void_p(s.cuda_stream)
        bd = C.byref(base.desc) if base is not None else None
        _lib.check(L.mcacq_acq_forward(C.byref(strat.desc), bd, C.byref(mc.desc), Xs[i].data_ptr(), b, q, acqs[i].data_ptr(),
                                       infos[i].data_ptr(), ws.data_ptr(), ws.numel(), sp), "fwd")
        if not fwd_only:
            _lib.check(L.mcacq_acq_backward(C.byref(strat.desc), bd, C.byref(mc.desc), Xs[i].data_ptr(), b, q, acqs[i].data_ptr(),
                                            ones.data_ptr(), gXs[i].data_ptr(), ws.data_ptr(), ws.numel(), sp), "bwd")
    for s in streams: torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / nchunks, acqs, gXs


ref = None
for ns in (1, 2, 3, 1, 2, 3):
    ms, acqs, gXs = run(ns)
    if ref is None: ref = (torch.stack(acqs).clone(), torch.stack(gXs).clone())
    same = bool(torch.equal(torch.stack(acqs), ref[0]) and torch.equal(torch.stack(gXs), ref[1]))
    print(f"{cfg} b={b} int8({strat.g_fwd},{strat.g_bwd}) fwd+bwd {ns} stream(s): {ms:.3f} ms per chunk   bit-identical to 1 stream: {same}", flush=True)
for ns in (1, 2, 1, 2):
    ms, _, _ = run(ns, fwd_only=True)
    print(f"{cfg} b={b} forward only {ns} stream(s): {ms:.3f} ms per chunk", flush=True)
