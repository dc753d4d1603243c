"""Developer tool: wall time vs kernel time of one optimiser round (b = num_restarts q-batches, forward + backward + the
D2H copies scipy needs) for C1 / C2 / C3."""
import os, sys, time
import torch
sys.path.insert(0, ".")
from botorch_b200.benchmarks import configs
from botorch_b200 import settings
from torch.profiler import profile, ProfilerActivity
settings.contraction.set(os.environ.get("MCACQ_CONTRACTION", "int8"))
dev = torch.device("cuda:0")
for cfg in ("C1", "C2", "C3"):
    spec = configs.CONFIGS[cfg]
    data = configs.make_problem(spec); model = configs.build_model(data, dev); acqf = configs.build_acqf(data, model)
    X = configs.eval_points(data, spec.num_restarts).to(dev)
    def round_():
        Xg = X.detach().requires_grad_(True); v = acqf(Xg); (g,) = torch.autograd.grad(v.sum(), Xg)
        return v.detach().cpu(), g.cpu()
    for _ in range(5): round_()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(50): round_()
    wall = (time.perf_counter() - t0) / 50 * 1e3
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(10): round_()
        torch.cuda.synchronize()
    kern = sum(e.device_time_total for e in prof.key_averages()) / 10e3
    nk = sum(e.count for e in prof.key_averages()) / 10
    print(f"{cfg} b={spec.num_restarts} n={spec.n}: wall {wall:.3f} ms per round, GPU busy {kern:.3f} ms in {nk:.0f} kernels/copies")
