"""Library marks for the FP64 roofline: torch.matmul (cuBLAS DGEMM) at 8192^3 and at the
tall-skinny shapes of the RSVD passes. Same method as MEASURED_PEAKS.json's bf16 entry."""
import json, sys, time
import torch

def best_ms(f, reps=10):
    f(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

out = {}
dev = "cuda:0"
N = 8192
a = torch.randn(N, N, dtype=torch.float64, device=dev); b = torch.randn(N, N, dtype=torch.float64, device=dev)
ms = best_ms(lambda: torch.matmul(a, b))
out["dgemm_8192_tflops_burst"] = 2 * N**3 / ms * 1e-9
t0 = time.time(); n = 0
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
while n < 40:
    torch.matmul(a, b); n += 1
e1.record(); torch.cuda.synchronize()
out["dgemm_8192_tflops_sustained"] = n * 2 * N**3 / e0.elapsed_time(e1) * 1e-9
del a, b
m, nn, l = 1 << 20, 1024, 112
A = torch.randn(m, nn, dtype=torch.float64, device=dev)
X = torch.randn(nn, l, dtype=torch.float64, device=dev)
Y = torch.randn(m, l, dtype=torch.float64, device=dev)
ms = best_ms(lambda: torch.matmul(A, X), 5)
out["cublas_AX_1Mx1024x112_tflops"] = 2 * m * nn * l / ms * 1e-9
ms = best_ms(lambda: torch.matmul(A.t(), Y), 5)
out["cublas_AtY_1Mx1024x112_tflops"] = 2 * m * nn * l / ms * 1e-9
ms = best_ms(lambda: torch.matmul(Y.t(), Y), 5)
out["cublas_YtY_1Mx112_tflops"] = 2 * m * l * l / ms * 1e-9
print(json.dumps(out, indent=1))
