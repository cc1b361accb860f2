"""cuBLAS DGEMM / TF32 GEMM throughput (library context for the K2 roofline)."""
import json
import torch

def bench(dtype, n, allow_tf32=False):
    torch.backends.cuda.matmul.allow_tf32 = allow_tf32
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    torch.matmul(a, b)
    best = 1e30
    for _ in range(6):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12

print(json.dumps({"cublas_dgemm_tflops_8192": round(bench(torch.float64, 8192), 2),
                  "cublas_tf32_tflops_8192": round(bench(torch.float32, 8192, True), 2),
                  "cublas_fp32_tflops_8192": round(bench(torch.float32, 8192, False), 2)}))
