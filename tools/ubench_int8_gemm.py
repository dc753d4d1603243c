"""Developer tool: what does the library reach on this part for int8 x int8 -> int32 (cuBLASLt through torch._int_mm) and
bf16 GEMMs?  The int8 figure is the practical denominator for the Ozaki kernel's tensor roofline."""
import torch
dev = torch.device("cuda:0")
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for n in (4096, 8192, 16384):
    a = torch.randint(-128, 127, (n, n), dtype=torch.int8, device=dev); b = torch.randint(-128, 127, (n, n), dtype=torch.int8, device=dev)
    try:
        ms = t(lambda: torch._int_mm(a, b.t()))
        print(f"int8 {n}^3 (A row-major, B^T view): {ms:.3f} ms  {2*n**3/ms*1e-9:.0f} TOP/s")
    except Exception as e:
        print("int_mm failed", str(e)[:200])
    try:
        ms = t(lambda: torch._int_mm(a, b))
        print(f"int8 {n}^3 (both row-major):          {ms:.3f} ms  {2*n**3/ms*1e-9:.0f} TOP/s")
    except Exception as e:
        print("int_mm failed", str(e)[:200])
    x = torch.randn(n, n, dtype=torch.bfloat16, device=dev); y = torch.randn(n, n, dtype=torch.bfloat16, device=dev)
    ms = t(lambda: x @ y)
    print(f"bf16 {n}^3: {ms:.3f} ms  {2*n**3/ms*1e-9:.0f} TFLOP/s")
