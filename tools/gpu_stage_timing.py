"""Per-stage device timings of the hot path at a BASELINE config (developer tool, run under gpurun)."""
import sys, time, ctypes as C, json
import torch
sys.path.insert(0, ".")
from botorch_b200 import _lib
from botorch_b200.models.prediction_strategy import DevicePredictionStrategy
from botorch_b200.acquisition._fused import BaselineOperands, MCOperands, fused_acquisition

dev = torch.device("cuda")
n, d, q, r, S = [int(x) for x in (sys.argv[1:6] if len(sys.argv) > 5 else (4096, 20, 8, 16, 1024))]
b = int(sys.argv[6]) if len(sys.argv) > 6 else 8192
kid = 1
torch.manual_seed(0)
X = torch.rand(n, d, dtype=torch.float64, device=dev)
Y = torch.sin(X.sum(-1) * 2) + 0.05 * torch.randn(n, dtype=torch.float64, device=dev)
ls = (0.5 + torch.rand(d, dtype=torch.float64, device=dev)) * 0.9
t0 = time.time()
strat = DevicePredictionStrategy(X, (Y - Y.mean()) / Y.std(), ls, torch.tensor(1e-3, device=dev, dtype=torch.float64), kid, 1.0, 0.0,
                                 torch.zeros(d, device=dev, dtype=torch.float64), torch.ones(d, device=dev, dtype=torch.float64), float(Y.mean()), float(Y.std()))
torch.cuda.synchronize(); print("setup s", time.time() - t0)
L = _lib.lib()
np_ = strat.np
M = b * q
f64 = dict(device=dev, dtype=torch.float64)
Xq = torch.rand(b, q, d, **f64)
U = strat.scale(Xq.view(-1, d))
Kt = torch.empty(M, np_, **f64); A = torch.empty(M, np_, **f64); counter = torch.zeros(64, dtype=torch.int32, device=dev)
st = _lib.stream_ptr()
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
t_cov = timeit(lambda: L.mcacq_cov_cross(kid, 1.0, U.data_ptr(), M, strat.U_train.data_ptr(), n, d, Kt.data_ptr(), np_, st))
print(f"cov_cross  {t_cov:.3f} ms  write {M*np_*8/t_cov*1e-6:.1f} GB/s")
for mode, name, Bm in ((0, "upper", strat.R), (1, "lower", strat.Rt), (2, "dense", strat.R)):
    t = timeit(lambda: L.mcacq_dgemm_tri(mode, M, np_, Kt.data_ptr(), Bm.data_ptr(), A.data_ptr(), counter.data_ptr(), st))
    fl = 2.0 * M * np_ * np_ * (0.5 if mode < 2 else 1.0)
    print(f"dgemm_tri {name:5s} {t:.3f} ms  {fl/t*1e-9:.2f} TF/s algorithmic ({2.0*M*np_*np_/t*1e-9:.2f} dense-equivalent)")
t = timeit(lambda: torch.matmul(Kt, strat.R, out=A))
print(f"cuBLAS dense dgemm {t:.3f} ms  {2.0*M*np_*np_/t*1e-9:.2f} TF/s")
# full fused forward / forward+backward
Xb = X[:r].contiguous()
Ub = strat.scale(Xb); Ab = strat.contracted_rows(Ub)
mean_b, cov_b = strat.posterior_blocks(Xb.unsqueeze(0))
Lb = torch.linalg.cholesky(cov_b[0])
Z = torch.randn(r + q, S, **f64)
best = torch.full((S,), float(Y.max()), **f64)
base = BaselineOperands(Ub, Ab, Lb.contiguous())
mc = MCOperands(Z, best, 1e-6, 1e-2, True)
def fwd():
    with torch.no_grad():
        return fused_acquisition(Xq, strat, base, mc)
t_f = timeit(fwd)
print(f"fused forward   {t_f:.3f} ms  {b*q*S/t_f*1e-6:.1f} Mpts/s")
def fwdbwd():
    Xg = Xq.detach().requires_grad_(True)
    v = fused_acquisition(Xg, strat, base, mc)
    torch.autograd.grad(v.sum(), Xg)
t_fb = timeit(fwdbwd)
print(f"fused fwd+bwd   {t_fb:.3f} ms  {b*q*S/t_fb*1e-6:.1f} Mpts/s   alg TF/s {(2.0*q*n*n*b)/t_fb*1e-9:.2f}")
