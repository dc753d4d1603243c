"""Developer tool: per-kernel CUDA-event-free timing of one fused forward+backward via torch.profiler (kernel name -> ms)."""
import os, sys
import torch
sys.path.insert(0, ".")
from botorch_b200.benchmarks import configs
from botorch_b200 import settings
settings.contraction.set(os.environ.get("MCACQ_CONTRACTION", "int8"))
cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
b = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
dev = torch.device("cuda:0")
from dataclasses import replace
spec = configs.CONFIGS[cfg]
if os.environ.get("MCACQ_R"): spec = replace(spec, r=int(os.environ["MCACQ_R"]))   # baseline size override
data = configs.make_problem(spec); model = configs.build_model(data, dev); acqf = configs.build_acqf(data, model)
X = configs.eval_points(data, b).to(dev)
def step():
    Xg = X.detach().requires_grad_(True); v = acqf(Xg); (g,) = torch.autograd.grad(v.sum(), Xg)
for _ in range(2): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 3e3) for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(t for _, t in rows)
for k, t in sorted(rows, key=lambda r: -r[1])[:12]:
    print(f"{t:8.3f} ms  {k[:90]}")
print(f"{tot:8.3f} ms  total kernel time per fwd+bwd (r = {spec.r})  [{os.environ.get('MCACQ_SR_NS','-')},{os.environ.get('MCACQ_SR_NSB','-')}]")
