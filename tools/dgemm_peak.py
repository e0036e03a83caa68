"""Measure cuBLAS DGEMM (torch.matmul fp64) burst + sustained TFLOP/s; the FP64 roofline denominator."""
import json, sys, time, torch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(3):
    torch.matmul(a, b)
torch.cuda.synchronize()
best = 1e30
for _ in range(10):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
burst = 2 * n**3 / best * 1e-9
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
t0 = time.time(); k = 0
e0.record()
while time.time() - t0 < 4.0:
    for _ in range(5):
        torch.matmul(a, b); k += 1
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
sus = 2 * n**3 * k / e0.elapsed_time(e1) * 1e-9
print(json.dumps({"dgemm_n": n, "fp64_tflops_burst": burst, "fp64_tflops_sustained": sus}))
