import sys, numpy as np, torch
sys.path.insert(0, ".")
from botorch_b200 import _lib
dev = torch.device("cuda:0"); L = _lib.lib(); st = _lib.stream_ptr(); f64 = dict(device=dev, dtype=torch.float64)
def np_slices(X, G):
    mx = np.abs(X).max(1); m, e = np.frexp(mx); shift = 8*G-2-e
    Xi = np.rint(np.ldexp(X, shift[:,None])).astype(np.int64)
    lim = (1 << (8*G-2)); Xi = np.clip(Xi, -lim, lim)
    S = np.zeros((G,)+X.shape, dtype=np.int64)
    for p in range(G-1,-1,-1):
        d = ((Xi & 0xFF) ^ 0x80) - 0x80; S[p] = d; Xi = (Xi - d) >> 8
    return S, np.ldexp(1.0, e+2)
def gpu_slices(X, G):
    rows, K = X.shape
    S = torch.empty(G, rows, K, dtype=torch.int8, device=dev); sc = torch.empty(rows, **f64)
    L.mcacq_slice_rows(X.data_ptr(), rows, K, K, K, G, 0, 0, S.data_ptr(), sc.data_ptr(), st); torch.cuda.synchronize()
    return S, sc
torch.manual_seed(0)
G = 6
for (M, K, N) in [(128, 64, 64), (128, 128, 64), (128, 192, 64), (128, 192, 192)]:
    A = torch.randn(M, K, **f64); B = torch.randn(N, K, **f64)
    As, ra = gpu_slices(A, G); Bs, cb = gpu_slices(B, G)
    nS, nr = np_slices(A.cpu().numpy(), G)
    print(f"M={M} K={K} N={N}: slices equal numpy: {np.array_equal(As.cpu().numpy().astype(np.int64), nS)}, scale equal: {np.array_equal(ra.cpu().numpy(), nr)}")
    C = torch.empty(M, N, **f64)
    L.mcacq_ozaki_contract(2, M, N, K, G, As.data_ptr(), ra.data_ptr(), Bs.data_ptr(), cb.data_ptr(), C.data_ptr(), N, st); torch.cuda.synchronize()
    ref = A @ B.t()
    # exact emulation with the GPU slices
    Asn = As.cpu().numpy().astype(np.int64); Bsn = Bs.cpu().numpy().astype(np.int64)
    Ce = np.zeros((M, N))
    for g in range(G):
        Ce += sum(Asn[p] @ Bsn[g-p].T for p in range(g+1)) * 256.0**-(g+2)
    Ce *= ra.cpu().numpy()[:,None] * cb.cpu().numpy()[None,:]
    err = (C - ref).abs() / ref.abs().max()
    print("   kernel vs fp64:", float(err.max()), " emulation-from-gpu-slices vs fp64:", float(np.abs(Ce - ref.cpu().numpy()).max()/ref.abs().max().item()))
    bad = (err > 1e-9).nonzero()
    print("   bad entries:", bad.shape[0], "rows:", sorted(set(bad[:,0].tolist()))[:10], "cols:", sorted(set(bad[:,1].tolist()))[:10])
