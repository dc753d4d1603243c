"""Developer tool (GPU): cProfile of optimize_acqf with the default (host, scipy `setulb`) driver -- where the host time of an
optimiser round goes."""
import cProfile, pstats, sys, time, warnings
import torch
sys.path.insert(0, ".")
from botorch_b200.benchmarks import configs
from botorch_b200.optim import optimize_acqf

dev = torch.device("cuda:0"); warnings.simplefilter("ignore")
cfg = sys.argv[1] if len(sys.argv) > 1 else "C1"
spec = configs.CONFIGS[cfg]
data = configs.make_problem(spec); model = configs.build_model(data, dev); acqf = configs.build_acqf(data, model)
bounds = torch.stack([torch.zeros(spec.d), torch.ones(spec.d)]).to(dev, torch.float64)
kw = dict(bounds=bounds, q=spec.q, num_restarts=spec.num_restarts, raw_samples=spec.raw_samples, options={"maxiter": 50, "seed": 0})
for _ in range(2):
    torch.manual_seed(0); optimize_acqf(acqf, **kw)
torch.cuda.synchronize(); t0 = time.perf_counter()
torch.manual_seed(0); optimize_acqf(acqf, **kw)
torch.cuda.synchronize(); print(f"{cfg}: optimize_acqf {1e3 * (time.perf_counter() - t0):.1f} ms")
pr = cProfile.Profile(); pr.enable()
torch.manual_seed(0); optimize_acqf(acqf, **kw)
torch.cuda.synchronize(); pr.disable()
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(45)
