import sys, time, torch
sys.path.insert(0, ".")
from cca_zoo_b200 import ops
torch.manual_seed(0)
for dt, n in [(torch.float64, 20000), (torch.float32, 20000)]:
    d = 1024
    v = [torch.randn(n, d, device="cuda", dtype=dt) for _ in range(2)]
    for _ in range(2): ops.moments(v, "exact")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): ops.moments(v, "exact")
    torch.cuda.synchronize(); t = (time.perf_counter() - t0) / 3
    F = n * 2 * d * (2 * d + 1)
    print(f"exact SIMT moments {dt} n={n} D={2*d}: {t*1e3:.2f} ms -> {F/t/1e12:.2f} TFLOP/s algorithmic")
A = torch.randn(2048, 2048, device="cuda", dtype=torch.float64); B = torch.randn(2048, 2048, device="cuda", dtype=torch.float64)
for _ in range(2): ops.gemm(A, B)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3): ops.gemm(A, B)
torch.cuda.synchronize(); t = (time.perf_counter() - t0) / 3
print(f"gemm f64 2048^3: {t*1e3:.2f} ms -> {2*2048**3/t/1e12:.2f} TFLOP/s")
