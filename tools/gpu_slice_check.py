"""Stage-by-stage GPU check of the kernels against the CPU oracle (developer tool, run under gpurun)."""
import sys, time, ctypes as C
import torch
sys.path.insert(0, ".")
from botorch_b200 import _lib
from botorch_b200.models.prediction_strategy import DevicePredictionStrategy
from botorch_b200.acquisition._fused import BaselineOperands, MCOperands, fused_acquisition
from oracle.gp import OracleGP
from oracle.acquisition import OracleQLogEI, OracleQLogNEI, value_and_grad
from oracle.sampling import qlogei_base_samples, qlognei_base_samples

torch.manual_seed(0)
dev = torch.device("cuda")

def rel(a, b):
    a = a.detach().cpu().double(); b = b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))

def run(n, d, q, r, S, b, kernel, outputscale, norm):
    X = torch.rand(n, d, dtype=torch.float64) * (3.0 if norm else 1.0)
    Y = torch.sin(X.sum(-1, keepdim=True)) + 0.05 * torch.randn(n, 1, dtype=torch.float64)
    ls = (0.5 + torch.rand(d, dtype=torch.float64)) * 0.9
    noise = torch.tensor(1e-3, dtype=torch.float64)
    off = torch.zeros(d, dtype=torch.float64); coef = torch.full((d,), 3.0 if norm else 1.0, dtype=torch.float64)
    gp = OracleGP(X, Y, ls, noise, kernel=kernel, outputscale=outputscale, mean_constant=0.1,
                  norm_offset=off if norm else None, norm_coef=coef if norm else None)
    Xt, Lc, mc_, cc, m, s = gp.caches()
    kid = 0 if kernel == "rbf" else 1
    t0 = time.time()
    strat = DevicePredictionStrategy(Xt.to(dev), ((Y - m) / s).squeeze(-1).to(dev), ls.to(dev), noise.to(dev), kid,
                                     1.0 if outputscale is None else outputscale, 0.1, off.to(dev), coef.to(dev), float(m), float(s))
    torch.cuda.synchronize()
    print(f"[n={n} d={d} q={q} r={r} S={S} b={b} {kernel}] setup {time.time()-t0:.3f}s  R rel err {rel(strat.R[:n,:n], cc):.2e} alpha {rel(strat.alpha[:n], mc_):.2e}")
    Xq = torch.rand(b, q, d, dtype=torch.float64) * (3.0 if norm else 1.0)
    mean_o, cov_o = gp.posterior_mvn(Xq)
    mean_g, cov_g = strat.posterior_blocks(Xq.to(dev))
    print(f"   posterior mean rel {rel(mean_g, mean_o):.2e}  covar rel {rel(cov_g, cov_o):.2e}  var rel(min elem) {float(((cov_g.cpu().diagonal(dim1=-1,dim2=-2)-cov_o.diagonal(dim1=-1,dim2=-2)).abs()/cov_o.diagonal(dim1=-1,dim2=-2)).max()):.2e}")
    if r == 0:
        best_f = float(Y.max())
        orc = OracleQLogEI(gp, torch.tensor(best_f, dtype=torch.float64), S, 1234)
        Z = qlogei_base_samples(S, q, 1234).view(S, q)
        best = torch.full((S,), best_f, dtype=torch.float64)
        base = None
    else:
        Xb = X[:r]
        orc = OracleQLogNEI(gp, Xb, S, 1234)
        Z = qlognei_base_samples(S, r, q, 1234).view(S, r + q)
        best = orc.baseline_best_f
        Ub = strat.scale(Xb.to(dev))
        Ab = strat.contracted_rows(Ub)
        base = BaselineOperands(Ub, Ab, orc.baseline_L.to(dev).contiguous())
    mc = MCOperands(Z.t().contiguous().to(dev), best.to(dev), 1e-6, 1e-2, True)
    v_o, g_o = value_and_grad(orc, Xq)
    Xg = Xq.to(dev).requires_grad_(True)
    v_g = fused_acquisition(Xg, strat, base, mc)
    (g_g,) = torch.autograd.grad(v_g.sum(), Xg)
    print(f"   acq rel {rel(v_g, v_o):.2e} (max abs {float((v_g.cpu()-v_o).abs().max()):.2e})  grad rel {rel(g_g, g_o):.2e}")
    print("   acq gpu", v_g[:4].tolist(), " oracle", v_o[:4].tolist())

run(64, 6, 4, 0, 512, 32, "rbf", None, False)
run(64, 6, 4, 5, 256, 16, "matern52", 1.3, True)
run(200, 20, 8, 16, 1024, 64, "rbf", None, True)
run(1000, 20, 8, 16, 1024, 256, "matern52", 1.0, True)
run(77, 3, 3, 7, 128, 9, "rbf", 2.0, False)
run(300, 5, 12, 20, 256, 10, "matern52", None, True)
