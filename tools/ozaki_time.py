"""Developer tool: time the int8 Ozaki contraction for a list of G (env MCACQ_OZ_BK forces the k-block bytes)."""
import sys, os
import torch
sys.path.insert(0, ".")
from botorch_b200 import _lib
dev = torch.device("cuda:0"); L = _lib.lib(); st = _lib.stream_ptr(); f64 = dict(device=dev, dtype=torch.float64)
M, n = 65536, 4096
A = torch.rand(M, n, **f64); R = torch.triu(torch.randn(n, n, **f64)); Rt = R.t().contiguous()
ref = A[:128] @ R
for G in [int(x) for x in sys.argv[1:]] or [4, 5, 6]:
    As = torch.empty(G, M, n, dtype=torch.int8, device=dev); ra = torch.empty(M, **f64)
    Bs = torch.empty(G, n, n, dtype=torch.int8, device=dev); cb = torch.empty(n, **f64)
    L.mcacq_slice_rows(A.data_ptr(), M, n, n, n, G, 1, 0, As.data_ptr(), ra.data_ptr(), st)
    L.mcacq_slice_rows(Rt.data_ptr(), n, n, n, n, G, 0, 0, Bs.data_ptr(), cb.data_ptr(), st)
    C = torch.empty(M, n, **f64)
    fn = lambda: L.mcacq_ozaki_contract(0, M, n, n, G, As.data_ptr(), ra.data_ptr(), Bs.data_ptr(), cb.data_ptr(), C.data_ptr(), n, st)
    rc = fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); fn(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    pairs = G * (G + 1) // 2
    err = float((C[:128] - ref).abs().max() / ref.abs().max())
    print(f"BK={os.environ.get('MCACQ_OZ_BK','auto')} G={G}: rc={rc} {ms:.3f} ms  int8 {M*n*(n+64.0)*pairs/ms*1e-12:.2f} POP/s  ms/G={ms/G:.3f}  err={err:.1e}")
