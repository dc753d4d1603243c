"""Golden vectors for the outcome-constraint helpers, generated from the REFERENCE's botorch/utils/objective.py and
botorch/utils/safe_math.py (build container only; same namespace shim as make_golden.py).
Usage: python tests/golden/make_golden_constraints.py"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import reference_modules  # noqa: E402


def main():
    sm, _, _ = reference_modules()
    import importlib

    ro = importlib.import_module("botorch.utils.objective")
    torch.manual_seed(3)
    x = torch.cat([torch.randn(200, dtype=torch.float64) * 30, torch.tensor([0.0, -1e3, 1e3, 17.9, 18.0, 18.1, 1e-9], dtype=torch.float64)])
    out = {"x": x, "logexpit": sm.logexpit(x), "log1pexp": sm.log1pexp(x), "fatmoid": sm.fatmoid(x, tau=0.7),
           "log_fatmoid": sm.log_fatmoid(x, tau=0.7), "sigmoid_fat_log": sm.sigmoid(x, log=True, fat=True),
           "sigmoid": sm.sigmoid(x)}
    samples = torch.randn(16, 3, 4, 2, dtype=torch.float64)
    cons = [lambda Y: Y[..., 1] - 0.2, lambda Y: -Y[..., 0] - 1.0]
    out["samples"] = samples
    for log in (False, True):
        for fat in (False, True):
            out[f"smoothed_log{int(log)}_fat{int(fat)}"] = ro.compute_smoothed_feasibility_indicator(
                cons, samples, eta=torch.tensor([1e-2, 0.5], dtype=torch.float64), log=log, fat=fat)
    out["smoothed_scalar_eta"] = ro.compute_smoothed_feasibility_indicator(cons, samples, eta=1e-3, log=True, fat=True)
    out["indicator"] = ro.compute_feasibility_indicator(cons, samples)
    out["indicator_marg"] = ro.compute_feasibility_indicator(cons, samples, marginalize_dim=-3)
    g = samples.clone().requires_grad_(True)
    val = ro.compute_smoothed_feasibility_indicator(cons, g, eta=1e-1, log=True, fat=True)
    out["smoothed_grad"] = torch.autograd.grad(val.sum(), g)[0]
    torch.save(out, os.path.join(HERE, "constraints_ref.pt"))
    print("wrote constraints_ref.pt", {k: tuple(v.shape) for k, v in out.items()})


if __name__ == "__main__":
    main()
