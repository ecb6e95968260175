"""Measures the FP64 denominators that MEASURED_PEAKS.json lacks: cuBLAS DGEMM (FP64 tensor
pipe) at 8192^3 and 4096^3, best of 10, CUDA events."""
import json
import torch

out = {}
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out["dgemm_%d_tflops" % n] = 2.0 * n ** 3 / (best * 1e-3) / 1e12
print(json.dumps(out))
