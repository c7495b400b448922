"""cuBLAS FP64 GEMM rate on this GPU (library reference point for the FP64 roofline)."""
import json, torch
n = 8192
a = torch.randn(n, n, device="cuda", dtype=torch.float64)
b = torch.randn(n, n, device="cuda", dtype=torch.float64)
for _ in range(2): torch.matmul(a, b)
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(json.dumps({"cublas_dgemm_tflops_8192": 2 * n**3 / best / 1e9, "ms": best}))
