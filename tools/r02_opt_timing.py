"""Developer tool (run under gpurun): wall time of optimize_acqf with the host (scipy setulb) and the device-resident
L-BFGS-B on C1-C3, and the per-round cost of each."""
import sys, time, warnings
import torch
sys.path.insert(0, ".")
from botorch_b200 import settings
from botorch_b200.benchmarks import configs
from botorch_b200.optim import optimize_acqf
from botorch_b200.generation import gen_candidates_device, gen_candidates_scipy

dev = torch.device("cuda:0")
warnings.simplefilter("ignore")
for name in sys.argv[1:] or ["C1", "C2", "C3"]:
    spec = configs.CONFIGS[name]
    data = configs.make_problem(spec)
    model = configs.build_model(data, dev)
    acqf = configs.build_acqf(data, model)
    bounds = torch.stack([torch.zeros(spec.d), torch.ones(spec.d)]).to(dev, torch.float64)
    kw = dict(bounds=bounds, q=spec.q, num_restarts=spec.num_restarts, raw_samples=spec.raw_samples, options={"maxiter": 50, "seed": 0})
    for mode in ("scipy", "device"):
        with settings.optimizer(mode):
            for rep in range(3):
                torch.manual_seed(0)
                torch.cuda.synchronize(); t0 = time.perf_counter()
                c, v = optimize_acqf(acqf, **kw)
                torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 1e3
            print(f"{name} optimize_acqf[{mode}]: {ms:.1f} ms  value {float(v):.10f}", flush=True)
    ics = configs.eval_points(data, spec.num_restarts, seed=5).to(dev)
    for fn, nm in ((gen_candidates_scipy, "scipy"), (gen_candidates_device, "device")):
        for rep in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            c, v = fn(ics, acqf, lower_bounds=0.0, upper_bounds=1.0, options={"maxiter": 50})
            torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 1e3
        extra = f" rounds {gen_candidates_device.last_rounds} -> {ms / max(1, gen_candidates_device.last_rounds):.3f} ms/round" if nm == "device" else ""
        print(f"{name} gen_candidates[{nm}]: {ms:.1f} ms  best {float(v.max()):.10f} mean {float(v.mean()):.8f}{extra}", flush=True)
