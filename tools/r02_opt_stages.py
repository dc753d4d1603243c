"""Developer tool (GPU): wall time of the stages of optimize_acqf on C3 (seeded): initial conditions (sweep + selection), the
optimiser, the rest."""
import sys, time, warnings
import torch
sys.path.insert(0, ".")
from botorch_b200 import settings
from botorch_b200.benchmarks import configs
from botorch_b200.optim import optimize as opt_mod
from botorch_b200.optim import optimize_acqf

dev = torch.device("cuda:0"); warnings.simplefilter("ignore")
cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
spec = configs.CONFIGS[cfg]
data = configs.make_problem(spec); model = configs.build_model(data, dev); acqf = configs.build_acqf(data, model)
bounds = torch.stack([torch.zeros(spec.d), torch.ones(spec.d)]).to(dev, torch.float64)
kw = dict(bounds=bounds, q=spec.q, num_restarts=spec.num_restarts, raw_samples=spec.raw_samples, options={"maxiter": 50, "seed": 0})
stages = {}
def wrap(mod, name):
    f = getattr(mod, name)
    def g(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = f(*a, **k)
        torch.cuda.synchronize(); stages[name] = stages.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
        return out
    setattr(mod, name, g)
import botorch_b200.generation.gen as gen_mod
import botorch_b200.generation.device_gen as dgen_mod
wrap(opt_mod, "gen_batch_initial_conditions")
wrap(gen_mod, "gen_candidates_scipy")
wrap(dgen_mod, "gen_candidates_device")
for mode in ("scipy", "device"):
    with settings.optimizer(mode):
        for rep in range(3):
            stages.clear()
            torch.manual_seed(0); torch.cuda.synchronize(); t0 = time.perf_counter()
            c, v = optimize_acqf(acqf, **kw)
            torch.cuda.synchronize(); total = (time.perf_counter() - t0) * 1e3
        extra = ""
        print(f"{cfg} [{mode}] total {total:.1f} ms: " + ", ".join(f"{k} {t:.1f}" for k, t in stages.items()) + f", rest {total - sum(stages.values()):.1f}{extra}  value {float(v):.10f}")
