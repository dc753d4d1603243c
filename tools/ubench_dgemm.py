import torch, time, json
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
def bench(M, N, K, reps=5):
    a = torch.randn(M, K, device=dev, dtype=torch.float64)
    b = torch.randn(K, N, device=dev, dtype=torch.float64)
    for _ in range(2): c = a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * M * N * K / best * 1e-9, best
for shp in [(8192, 8192, 8192), (65536, 4096, 4096), (16384, 1024, 1024), (4096, 4096, 4096)]:
    tf, ms = bench(*shp)
    print("cuBLAS DGEMM", shp, "%.2f TF/s" % tf, "%.2f ms" % ms, flush=True)
# sustained
a = torch.randn(8192, 8192, device=dev, dtype=torch.float64); b = torch.randn(8192, 8192, device=dev, dtype=torch.float64)
torch.cuda.synchronize(); t0 = time.time(); n = 0
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < 4.0:
    for _ in range(5): c = a @ b
    n += 5; torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
print("cuBLAS DGEMM sustained 8192^3: %.2f TF/s" % (2.0 * 8192**3 * n / e0.elapsed_time(e1) * 1e-9))
