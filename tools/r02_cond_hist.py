"""Developer diagnostic (GPU): distribution of the per-q-batch conditioning byte floor(-4 log2 rho) on the benchmark
configurations -- the Sobol sweep points and the candidates an optimize_acqf run ends at -- to calibrate
DevicePredictionStrategy.INT8_COND_LIMIT."""
import os, sys, warnings
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from botorch_b200 import _lib, settings
from botorch_b200.acquisition._fused import FusedMCAcquisition, RerouteStats
from botorch_b200.benchmarks import configs
from botorch_b200.optim import optimize_acqf

dev = torch.device("cuda:0")
warnings.simplefilter("ignore")
for cfg in ("C1", "C2", "C3"):
    spec = configs.CONFIGS[cfg]
    data = configs.make_problem(spec)
    model = configs.build_model(data, dev)
    acqf = configs.build_acqf(data, model)
    strat = model.prediction_strategy()
    X = configs.eval_points(data, min(spec.raw_samples, 8192)).to(dev)
    base = acqf._baseline_operands() if hasattr(acqf, "_baseline_operands") else None
    with torch.no_grad():
        _, info = FusedMCAcquisition.apply(X, strat, base, acqf._mc_operands(X))
    cond = ((info >> 8) & 0xFF).cpu()
    qs = torch.quantile(cond.double(), torch.tensor([0.0, 0.5, 0.9, 0.99, 0.999, 1.0], dtype=torch.float64)).tolist()
    print(f"{cfg}: mode {strat.contraction}({strat.g_fwd},{strat.g_bwd}) sweep cond byte min/50/90/99/99.9/max = {qs}  (rho_min = {2 ** (-qs[-1] / 4):.3e})")
    vb = ((info >> 16) & 0xFF).cpu()
    qv = torch.quantile(vb.double(), torch.tensor([0.0, 0.5, 0.9, 0.99, 0.999, 1.0], dtype=torch.float64)).tolist()
    print(f"    variance-collapse byte min/50/90/99/99.9/max = {qv} (smallest var/prior = {2 ** (-qv[-1] / 4):.3e}); model limit: ratio {strat.int8_var_ratio_limit} byte {strat.int8_var_byte_limit}; probe err {strat.int8_probe_error:.2e} grad {strat.int8_probe_grad_error:.2e}")
    ob = torch.stack([torch.zeros(spec.d), torch.ones(spec.d)]).to(dev, torch.float64)
    RerouteStats.q_batches = RerouteStats.calls = 0
    cand, val = optimize_acqf(acqf, bounds=ob, q=spec.q, num_restarts=spec.num_restarts, raw_samples=min(spec.raw_samples, 8192),
                              options={"maxiter": 50, "seed": 0}, return_best_only=False)
    with torch.no_grad():
        _, info = FusedMCAcquisition.apply(cand.contiguous(), strat, base, acqf._mc_operands(cand))
    cond = ((info >> 8) & 0xFF).cpu()
    print(f"    optimised candidates ({cand.shape[0]} restarts): cond byte min/median/max = {int(cond.min())}/{int(cond.median())}/{int(cond.max())};"
          f" rerouted during optimize_acqf: {RerouteStats.q_batches} q-batches in {RerouteStats.calls} calls")
