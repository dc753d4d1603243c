"""Feasibility prototype (developer tool): emulate the fp64 contraction A = Kt R with int8 tensor-core GEMMs
(Ozaki-style slicing, torch._int_mm) and measure accuracy of the posterior covariance and the int8 GEMM rate."""
import sys, time
import torch
sys.path.insert(0, ".")
from dataclasses import replace
from botorch_b200.benchmarks import configs

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
b = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
data = configs.make_problem(replace(configs.C3, n=n))
model = configs.build_model(data, dev)
strat = model.prediction_strategy()
X = configs.eval_points(data, b).to(dev)
q = data.spec.q
U = strat.scale(X.view(-1, X.shape[-1]))
M = U.shape[0]
f64 = dict(device=dev, dtype=torch.float64)
Kt = torch.empty(M, strat.np, **f64)
from botorch_b200 import _lib
L = _lib.lib()
L.mcacq_cov_cross(strat.kernel_id, strat.outputscale, U.data_ptr(), M, strat.U_train.data_ptr(), strat.n, strat.d, Kt.data_ptr(), strat.np, _lib.stream_ptr())
R = strat.R
A_ref = Kt @ R

def slice_rows(Mx, s):
    """row-scaled 7-bit slices: Mx[i,:] = 2^e_i * sum_p q_p[i,:] 128^-(p+1)"""
    mx = Mx.abs().amax(dim=1, keepdim=True).clamp_min(1e-300)
    e = torch.ceil(torch.log2(mx)) + 1e-9
    e = torch.ceil(torch.log2(mx))
    r = Mx * torch.exp2(-e)
    # guard: |r| <= 1; make strictly < 1 by bumping exponent where equal
    bump = (r.abs().amax(dim=1, keepdim=True) >= 1.0)
    e = e + bump.to(e.dtype); r = Mx * torch.exp2(-e)
    out = []
    for p in range(s):
        r = r * 128.0
        qv = torch.trunc(r)
        out.append(qv.to(torch.int8))
        r = r - qv
    return out, e

for s in (5, 6, 7, 8):
    As, ea = slice_rows(Kt, s)
    Bs_t, eb = slice_rows(R.t().contiguous(), s)   # column scaling of R == row scaling of R^T
    Bs = [x.t().contiguous() for x in Bs_t]          # K x N int8
    torch.cuda.synchronize(); t0 = time.time()
    C = torch.zeros(M, strat.np, **f64)
    for g in range(s):
        Acat = torch.cat([As[p] for p in range(g + 1)], dim=1)                 # M x (g+1)K
        Bcat = torch.cat([Bs[g - p] for p in range(g + 1)], dim=0)             # (g+1)K x N
        Sg = torch._int_mm(Acat, Bcat)                                          # int32
        C += Sg.to(torch.float64) * (128.0 ** (-(g + 2)))
    C = C * torch.exp2(ea) * torch.exp2(eb).t()
    torch.cuda.synchronize(); t_all = time.time() - t0
    err = (C - A_ref).abs().max() / A_ref.abs().max()
    # downstream: posterior variance per point  var = k(x,x) - |A_row|^2
    var_ref = strat.outputscale - (A_ref * A_ref).sum(1)
    var_emu = strat.outputscale - (C * C).sum(1)
    relvar = ((var_emu - var_ref).abs() / var_ref.abs()).max()
    print(f"s={s}: max rel err A {float(err):.2e}  posterior-variance rel err {float(relvar):.2e}  (python-loop time {t_all*1e3:.1f} ms)")

# raw int8 GEMM rate
for kmul in (1, 3, 6):
    a = torch.randint(-127, 127, (M, kmul * strat.np), device=dev, dtype=torch.int8)
    bm = torch.randint(-127, 127, (kmul * strat.np, strat.np), device=dev, dtype=torch.int8)
    torch._int_mm(a, bm); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); 
    for _ in range(3): torch._int_mm(a, bm)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"int8 GEMM M={M} K={kmul*strat.np} N={strat.np}: {ms:.3f} ms  {2.0*M*kmul*strat.np*strat.np/ms*1e-9:.0f} TOP/s")
t = A_ref  # fp64 reference time
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); Kt @ R; e1.record(); torch.cuda.synchronize(); print("cuBLAS fp64 dense: %.3f ms" % e0.elapsed_time(e1))
