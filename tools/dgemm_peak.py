"""Measure cuBLAS DGEMM throughput (the fp64 roofline denominator).  Prints one JSON line."""
import json, time, torch
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
c = torch.empty_like(a)
for _ in range(3):
    torch.matmul(a, b, out=c)
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
burst = 2 * n**3 / best * 1e-9
# sustained: back-to-back for ~4 s
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
t0 = time.time(); k = 0
e0.record()
while time.time() - t0 < 4.0:
    for _ in range(5):
        torch.matmul(a, b, out=c); k += 1
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
sust = 2 * n**3 * k / e0.elapsed_time(e1) * 1e-9
print(json.dumps({"fp64_dgemm_tflops": round(burst, 2), "fp64_dgemm_tflops_sustained": round(sust, 2),
                  "n": n, "how": "torch.matmul float64 8192^3 (cuBLAS): best of 10 (burst), 4 s back-to-back (sustained)"}))
