"""Developer tool: correctness and speed of the tcgen05 int8 Ozaki contraction against fp64 matmul (run under gpurun)."""
import sys, time
import torch
sys.path.insert(0, ".")
from botorch_b200 import _lib

dev = torch.device("cuda:0")
L = _lib.lib()
st = _lib.stream_ptr()
f64 = dict(device=dev, dtype=torch.float64)

def slice_rows(X, G, Kp=None, fixed=None):
    rows, K = X.shape
    Kp = Kp or K
    S = torch.empty(G, rows, Kp, dtype=torch.int8, device=dev)
    sc = torch.empty(rows, **f64)
    _lib.check(L.mcacq_slice_rows(X.data_ptr(), rows, K, X.stride(0), Kp, G, int(fixed is not None), fixed or 0, S.data_ptr(), sc.data_ptr(), st), "slice")
    return S, sc

def contract(mode, A, Bt, G):
    M, K = A.shape; N = Bt.shape[0]
    As, ra = slice_rows(A, G); Bs, cb = slice_rows(Bt, G)
    C = torch.empty(M, N, **f64)
    _lib.check(L.mcacq_ozaki_contract(mode, M, N, K, G, As.data_ptr(), ra.data_ptr(), Bs.data_ptr(), cb.data_ptr(), C.data_ptr(), N, st), "contract")
    torch.cuda.synchronize()
    return C, (As, ra, Bs, cb)

torch.manual_seed(0)
for (M, n, mode) in [(128, 64, 2), (300, 192, 2), (1000, 1024, 0), (1000, 1024, 1), (777, 1040, 0), (4096, 4096, 0)]:
    A = torch.randn(M, n, **f64) * torch.exp(torch.randn(M, 1, **f64))
    R = torch.randn(n, n, **f64)
    if mode == 0: R = torch.triu(R)
    if mode == 1: R = torch.tril(R)   # B[k][j] nonzero for k >= j
    ref = A @ R
    for G in (4, 5, 6):
        C, _ = contract(mode, A, R.t().contiguous(), G)
        err = float((C - ref).abs().max() / ref.abs().max())
        print(f"M={M} n={n} mode={mode} G={G}: max rel err {err:.2e}")

# timing at the C3 chunk size
M, n = 65536, 4096
A = torch.rand(M, n, **f64)
R = torch.triu(torch.randn(n, n, **f64))
for G in (4, 5, 6):
    As, ra = slice_rows(A, G, fixed=0)
    Bs, cb = slice_rows(R.t().contiguous(), G)
    C = torch.empty(M, n, **f64)
    for mode, name in ((0, "upper"), (2, "dense")):
        fn = lambda: L.mcacq_ozaki_contract(mode, M, n, n, G, As.data_ptr(), ra.data_ptr(), Bs.data_ptr(), cb.data_ptr(), C.data_ptr(), n, st)
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); fn(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
        pairs = G * (G + 1) // 2
        frac = 0.5 if mode == 0 else 1.0
        print(f"G={G} {name}: {ms:.3f} ms  -> {2.0*M*n*n*frac/ms*1e-9:.1f} TF/s fp64-equivalent, int8 rate {2.0*M*n*n*frac*pairs/ms*1e-12:.2f} POP/s")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); slice_rows(A, G, fixed=0); e1.record(); torch.cuda.synchronize()
    print(f"   slicing A (M x n fp64 -> {G} int8 slices): {e0.elapsed_time(e1):.3f} ms")
ref = A[:256] @ R
print("spot check err", float((C[:256] - ref).abs().max() / ref.abs().max()))
