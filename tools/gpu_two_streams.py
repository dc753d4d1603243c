"""Developer experiment: do the SIMT kernels of one chunk hide under the int8 contraction of another when consecutive
chunks run on two streams?  Prints ms per chunk for 1 and 2 streams."""
import os, sys
import torch
sys.path.insert(0, ".")
from botorch_b200.benchmarks import configs
from botorch_b200 import settings
settings.contraction.set(os.environ.get("MCACQ_CONTRACTION", "int8"))
dev = torch.device("cuda:0")
data = configs.make_problem(configs.CONFIGS["C3"]); model = configs.build_model(data, dev); acqf = configs.build_acqf(data, model)
nchunks, b = 8, 8192
Xs = [configs.eval_points(data, b).to(dev) + 1e-3 * i for i in range(nchunks)]
def one(X):
    Xg = X.detach().requires_grad_(True); v = acqf(Xg); (g,) = torch.autograd.grad(v.sum(), Xg); return g
def run(nstreams):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams: s.wait_stream(torch.cuda.current_stream())
    outs = []
    for i, X in enumerate(Xs):
        with torch.cuda.stream(streams[i % nstreams]):
            outs.append(one(X))
    for s in streams: torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / nchunks, outs
for _ in range(2): one(Xs[0])
for ns in (1, 2, 3, 1, 2):
    ms, outs = run(ns)
    print(f"{ns} stream(s): {ms:.3f} ms per chunk")
