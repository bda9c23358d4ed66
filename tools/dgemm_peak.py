"""cuBLAS DGEMM calibration (FP64 roofline denominator), run on the B200 box."""
import json, torch, time
torch.backends.cuda.matmul.allow_tf32 = False
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
c = torch.empty_like(a)
for _ in range(3):
    torch.matmul(a, b, out=c)
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
burst = 2 * n**3 / best / 1e9
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); k = 0
t0 = time.time()
while time.time() - t0 < 4:
    for _ in range(5):
        torch.matmul(a, b, out=c); k += 1
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
sustained = 2 * n**3 * k / e0.elapsed_time(e1) / 1e9
print(json.dumps({"dgemm_tflops_burst": burst, "dgemm_tflops_sustained": sustained, "n": n, "best_ms": best}))
