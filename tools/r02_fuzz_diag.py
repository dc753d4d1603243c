"""Developer diagnostic (GPU): for given fuzz seeds, the value / gradient error of the fused path against the oracle for the
auto-selected slice counts, for pinned (g_fwd, g_bwd) pairs and for 'dmma', next to the conditioning of the joint covariance
(smallest Cholesky pivot relative to the prior variance).  Usage: python tools/r02_fuzz_diag.py 28 70"""
import os
import sys
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
DEV = torch.device("cuda:0")


def run(seed, mode, slices):
    from test_gpu_fuzz import _case

    from botorch_b200 import settings
    from botorch_b200.acquisition import qLogExpectedImprovement, qLogNoisyExpectedImprovement
    from botorch_b200.models import MaternKernel, RBFKernel, ScaleKernel, SingleTaskGP
    from botorch_b200.models.transforms import Normalize
    from botorch_b200.sampling import SobolQMCNormalSampler
    from oracle.acquisition import OracleQLogEI, OracleQLogNEI, value_and_grad
    from oracle.gp import OracleGP, psd_safe_cholesky

    c = _case(seed)
    g, n, d, q, r, S, b = c["g"], c["n"], c["d"], c["q"], c["r"], c["S"], c["b"]
    lo, hi = (-2.0, 3.0) if c["normalize"] else (0.0, 1.0)
    X = lo + (hi - lo) * torch.rand(n, d, generator=g, dtype=torch.float64)
    Y = torch.sin(2.5 * (X - lo).sum(-1, keepdim=True) / (hi - lo) / d ** 0.5) + 0.1 * torch.randn(n, 1, generator=g, dtype=torch.float64)
    ls = (0.15 + 0.25 * torch.rand(d, generator=g, dtype=torch.float64)) * d ** 0.5
    os_ = 1.7 if c["scale"] else None
    bounds = torch.tensor([[lo] * d, [hi] * d], dtype=torch.float64)
    noise = (2e-3 + 5e-3 * torch.rand(n, generator=g, dtype=torch.float64)) if c["fixed_noise"] else torch.tensor(4e-3, dtype=torch.float64)
    base = (RBFKernel if c["kernel"] == "rbf" else MaternKernel)(ard_num_dims=d, lengthscale=ls)
    with settings.contraction(mode), settings.int8_slices(slices), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        s_y = float(Y.std()) if n > 1 else 1.0
        model = SingleTaskGP(X.to(DEV), Y.to(DEV),
                             train_Yvar=(noise * s_y**2).unsqueeze(-1).to(DEV) if c["fixed_noise"] else None,
                             covar_module=ScaleKernel(base, os_) if os_ else base,
                             input_transform=Normalize(d=d, bounds=bounds.to(DEV)) if c["normalize"] else None).to(DEV)
        if not c["fixed_noise"]:
            model.likelihood.noise = float(noise)
        gp = OracleGP(X, Y, ls, noise, kernel=c["kernel"], outputscale=os_,
                      norm_offset=bounds[0] if c["normalize"] else None,
                      norm_coef=(bounds[1] - bounds[0]) if c["normalize"] else None)
        Xq = lo + (hi - lo) * torch.rand(b, q, d, generator=g, dtype=torch.float64)
        sampler = SobolQMCNormalSampler(torch.Size([S]), seed=seed)
        if r == 0:
            best = Y.max() - 0.2
            acqf = qLogExpectedImprovement(model, best_f=best.to(DEV), sampler=sampler, fat=c["fat"])
            orc = OracleQLogEI(gp, best, S, seed, fat=c["fat"])
            Xall = Xq
        else:
            Xb = lo + (hi - lo) * torch.rand(r, d, generator=g, dtype=torch.float64)
            acqf = qLogNoisyExpectedImprovement(model, X_baseline=Xb.to(DEV), prune_baseline=False, sampler=sampler, fat=c["fat"])
            orc = OracleQLogNEI(gp, Xb, S, seed, fat=c["fat"])
            Xall = torch.cat([Xb.expand(b, r, d), Xq], dim=-2)
        _, cov = gp.posterior_mvn(Xall)
        L = psd_safe_cholesky(cov, max_tries=6)
        prior = float(cov.diagonal(dim1=-1, dim2=-2).max())
        piv = L.diagonal(dim1=-1, dim2=-2) ** 2 / prior
        piv_base = float(piv[:, :r].min()) if r else float("nan")
        piv_q = float(piv[:, r:].min())
        v_o, g_o = value_and_grad(orc, Xq)
        Xg = Xq.to(DEV).requires_grad_(True)
        v = acqf(Xg)
        (gr,) = torch.autograd.grad(v.sum(), Xg)
        st = model.prediction_strategy()
        ev = float(((v.detach().cpu() - v_o).abs() / v_o.abs().clamp_min(1e-12)).max())
        eg = float((gr.cpu() - g_o).abs().max() / g_o.abs().max().clamp_min(1e-300))
        tag = f"{st.contraction}({st.g_fwd},{st.g_bwd})" if st.contraction == "int8" else "dmma"
        print(f"  seed {seed:3d} n={n} d={d} q={q} r={r} S={S} b={b} {c['kernel']:8s} want {mode}{slices or ''} -> {tag:10s} "
              f"probe v={st.int8_probe_error} g={st.int8_probe_grad_error}  pivot base {piv_base:.2e} q {piv_q:.2e}  "
              f"value err {ev:.2e} grad err {eg:.2e}", flush=True)


if __name__ == "__main__":
    seeds = [int(a) for a in sys.argv[1:]] or [28, 70]
    for s in seeds:
        for mode, sl in (("dmma", None), ("int8", None), ("int8", (6, 5)), ("int8", (7, 5)), ("int8", (7, 6)), ("int8", (7, 7))):
            try:
                run(s, mode, sl)
            except Exception as e:  # noqa: BLE001
                print(f"  seed {s} {mode} {sl}: {type(e).__name__}: {e}")
