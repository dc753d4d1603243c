"""Two fused forward+backward passes at a BASELINE config (profiling target; run under ncu via gpurun)."""
import sys
import torch
sys.path.insert(0, ".")
from dataclasses import replace
from botorch_b200.benchmarks import configs

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
b = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda:0")
import os
from botorch_b200 import settings
settings.contraction.set(os.environ.get("MCACQ_CONTRACTION", "int8"))
if os.environ.get("MCACQ_SLICES"):
    settings.int8_slices.set(tuple(int(v) for v in os.environ["MCACQ_SLICES"].split(",")))
data = configs.make_problem(configs.CONFIGS[cfg])
model = configs.build_model(data, dev)
acqf = configs.build_acqf(data, model)
X = configs.eval_points(data, b).to(dev)
for _ in range(reps):
    Xg = X.detach().requires_grad_(True)
    v = acqf(Xg)
    (g,) = torch.autograd.grad(v.sum(), Xg)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
prof_last = os.environ.get("MCACQ_PROFILE_LAST") == "1"   # ncu --profile-from-start off: only the timed pass is captured
if prof_last: torch.cuda.profiler.start()
e0.record()
Xg = X.detach().requires_grad_(True)
v = acqf(Xg)
(g,) = torch.autograd.grad(v.sum(), Xg)
e1.record(); torch.cuda.synchronize()
if prof_last: torch.cuda.profiler.stop()
print(f"{cfg} b={b}: fwd+bwd {e0.elapsed_time(e1):.3f} ms")
