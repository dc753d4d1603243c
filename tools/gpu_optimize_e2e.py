"""Developer tool: wall-clock anatomy of `optimize_acqf` (raw-sample sweep + batched L-BFGS-B rounds) on C1 / C2 / C3."""
import os, sys, time
import torch
sys.path.insert(0, ".")
from botorch_b200.benchmarks import configs
from botorch_b200 import settings
from botorch_b200.optim import optimize_acqf
from botorch_b200.acquisition._fused import LaunchStats
from botorch_b200.generation import gen as _gen
_orig_fmin = _gen.fmin_l_bfgs_b_batched
_last = {}
def _capture(*a, **k):
    out = _orig_fmin(*a, **k); _last['res'] = out[2]; return out
_gen.fmin_l_bfgs_b_batched = _capture
settings.contraction.set(os.environ.get("MCACQ_CONTRACTION", "int8"))
dev = torch.device("cuda:0")
for cfg in sys.argv[1:] or ["C1", "C2"]:
    spec = configs.CONFIGS[cfg]
    data = configs.make_problem(spec); model = configs.build_model(data, dev); acqf = configs.build_acqf(data, model)
    bounds = torch.stack([torch.zeros(spec.d), torch.ones(spec.d)]).to(dev, torch.float64)
    calls = {"n": 0, "t": 0.0}
    orig = type(acqf).forward
    def timed_forward(self, X, _o=orig):
        torch.cuda.synchronize(); t0 = time.perf_counter(); out = _o(self, X); torch.cuda.synchronize()
        calls["n"] += 1; calls["t"] += time.perf_counter() - t0; return out
    type(acqf).forward = timed_forward
    for rep in range(2):
        calls.update(n=0, t=0.0); torch.cuda.synchronize(); t0 = time.perf_counter()
        cand, val = optimize_acqf(acqf, bounds=bounds, q=spec.q, num_restarts=spec.num_restarts, raw_samples=spec.raw_samples,
                                  options={"maxiter": 50, "seed": 0})
        torch.cuda.synchronize(); wall = time.perf_counter() - t0
    type(acqf).forward = orig
    import collections
    res = _last.get("res", [])
    print("   nit:", sorted(r.nit for r in res)[::max(1, len(res)//8)], "status:", dict(collections.Counter((r.status, str(r.message)[:14]) for r in res)))
    print(f"{cfg}: optimize_acqf wall {wall*1e3:.1f} ms; {calls['n']} acqf calls, {calls['t']*1e3:.1f} ms inside forward (fwd only, synced); value {float(val):.6f}")
