"""Generate golden vectors from the REFERENCE's own modules (run in the build container only).

Imports the gpytorch-free leaf modules of /root/reference/botorch through a namespace shim that bypasses
botorch/__init__.py (which needs gpytorch): utils/safe_math.py, sampling/qmc.py, utils/sampling.py.  Writes small
.pt fixtures next to this script.  The GP part of the path lives in un-vendored gpytorch/linear_operator, so no
reference-generated goldens exist for it ("parity unpinned", oracle/__init__.py); `oracle_path_*.pt` are minted from
the oracle restatement itself and only guard against regressions of the oracle.
Usage: python tests/golden/make_golden.py
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def reference_modules():
    for name, sub in (("botorch", ""), ("botorch.utils", "utils"), ("botorch.sampling", "sampling")):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            mod.__path__ = [os.path.join(REF, "botorch", sub)]
            sys.modules[name] = mod
    import importlib

    sm = importlib.import_module("botorch.utils.safe_math")
    qmc = importlib.import_module("botorch.sampling.qmc")
    us = importlib.import_module("botorch.utils.sampling")
    return sm, qmc, us


def main():
    sm, qmc, us = reference_modules()
    torch.manual_seed(0)
    # ---- reduction chain: log_fatplus -> fatmax -> logmeanexp on realistic improvement values
    S, b, q = 64, 5, 4
    obj = torch.randn(S, b, q, dtype=torch.float64) * 0.7
    obj[0, 0] = torch.tensor([3.0, -2.0, 0.0, 1e-7], dtype=torch.float64)
    best = torch.tensor(0.3, dtype=torch.float64)
    x = (obj - best).clone().requires_grad_(True)
    li = sm.log_fatplus(x, tau=1e-6)
    fm = sm.fatmax(li, dim=-1, tau=1e-2)
    acq = sm.logmeanexp(fm, dim=0)
    (g,) = torch.autograd.grad(acq.sum(), x)
    x2 = (obj - best).clone().requires_grad_(True)
    li2 = sm.log_softplus(x2, tau=1e-6)
    fm2 = sm.smooth_amax(li2, dim=-1, tau=1e-2)
    acq2 = sm.logmeanexp(fm2, dim=0)
    (g2,) = torch.autograd.grad(acq2.sum(), x2)
    torch.save({"z": (obj - best), "li": li.detach(), "fatmax": fm.detach(), "acq": acq.detach(), "grad": g,
                "li_nofat": li2.detach(), "smooth_amax": fm2.detach(), "acq_nofat": acq2.detach(), "grad_nofat": g2,
                "tau_relu": 1e-6, "tau_max": 1e-2}, os.path.join(HERE, "safe_math_chain.pt"))
    # ---- inf handling of the helpers
    xi = torch.tensor([[0.0, float("inf"), 1.0], [float("-inf")] * 3, [1.0, 2.0, 3.0]], dtype=torch.float64)
    torch.save({"x": xi, "fatmax": sm.fatmax(xi, dim=-1, tau=1e-2), "logsumexp": sm.logsumexp(xi, dim=-1),
                "logmeanexp": sm.logmeanexp(xi, dim=-1)}, os.path.join(HERE, "safe_math_inf.pt"))
    # ---- qMC normal base samples and Sobol X draws
    draws = {}
    for (d, n, seed) in [(4, 16, 1234), (24, 32, 1234), (8, 8, 7)]:
        draws[f"normal_d{d}_n{n}_s{seed}"] = us.draw_sobol_normal_samples(d=d, n=n, dtype=torch.float64, seed=seed)
    bounds = torch.stack([torch.zeros(6, dtype=torch.float64), torch.ones(6, dtype=torch.float64)])
    draws["sobol_n8_q4_d6_s0"] = us.draw_sobol_samples(bounds=bounds, n=8, q=4, seed=0)
    bounds2 = torch.tensor([[-1.0, 0.0, 2.0], [1.0, 5.0, 2.5]], dtype=torch.float64)
    draws["sobol_n4_q2_d3_s11"] = us.draw_sobol_samples(bounds=bounds2, n=4, q=2, seed=11)
    w = torch.tensor([0.3, -1.2, 0.8, 2.5, 2.4, -0.1], dtype=torch.float64)
    with us.manual_seed(5):
        draws["boltzmann_idx"] = us.boltzmann_sample(w, num_samples=3, eta=2.0)
    torch.save(draws, os.path.join(HERE, "qmc_draws.pt"))
    # ---- oracle-minted regression vectors for the GP part (NOT reference outputs)
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle.acquisition import OracleQLogEI, OracleQLogNEI, value_and_grad
    from oracle.gp import OracleGP

    torch.manual_seed(1)
    n, d, qq = 24, 3, 2
    X = torch.rand(n, d, dtype=torch.float64)
    Y = torch.sin(3 * X.sum(-1, keepdim=True)) + 0.05 * torch.randn(n, 1, dtype=torch.float64)
    ls = torch.tensor([0.4, 0.7, 0.55], dtype=torch.float64)
    Xq = torch.rand(3, qq, d, dtype=torch.float64)
    out = {"X": X, "Y": Y, "ls": ls, "Xq": Xq}
    for kern in ("rbf", "matern52"):
        gp = OracleGP(X, Y, ls, torch.tensor(1e-3, dtype=torch.float64), kernel=kern, outputscale=1.7 if kern == "matern52" else None)
        mean, cov = gp.posterior_mvn(Xq)
        ei = OracleQLogEI(gp, Y.max(), 64, 1234)
        nei = OracleQLogNEI(gp, X[:5], 64, 1234)
        v1, g1 = value_and_grad(ei, Xq)
        v2, g2 = value_and_grad(nei, Xq)
        out[kern] = {"mean": mean, "cov": cov, "qlogei": v1, "qlogei_grad": g1, "qlognei": v2, "qlognei_grad": g2}
    torch.save(out, os.path.join(HERE, "oracle_path_small.pt"))
    # ---- REAL reference outputs of the compiled botorch/csrc/logei_fused.cpp (oracle/build_ref.py)
    from oracle.build_ref import build, load_ref

    build()
    ref = load_ref()
    torch.manual_seed(2)
    cases = {}
    for name, (B, n_sub, isz, m, nc, batched) in {"ehvi": (6, 5, 3, 2, 7, False), "nehvi": (4, 6, 2, 3, 5, True),
                                                   "single": (3, 4, 1, 4, 3, False), "wide": (2, 3, 9, 2, 4, True)}.items():
        obj = torch.randn(B, n_sub, isz, m, dtype=torch.float64)
        lo = torch.randn(*( (B,) if batched else () ), nc, m, dtype=torch.float64) - 0.5
        hi = lo + torch.rand_like(lo) * 2 + 0.05
        hi[..., -1, :] = float("inf")  # unbounded top cell, clamped to 1e10 inside the kernel
        go = torch.randn(B, nc, n_sub, dtype=torch.float64)
        fwd = ref.forward(obj, lo, hi, 1e-6, 1e-2)
        bwd = ref.backward(go, obj, lo, hi, 1e-6, 1e-2)
        cases[name] = {"obj": obj, "lo": lo, "hi": hi, "go": go, "fwd": fwd, "bwd": bwd}
    torch.save(cases, os.path.join(HERE, "log_areas_ref.pt"))
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
