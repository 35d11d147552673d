"""Time the individual kernels on the BASELINE shapes (CUDA events, after warm-up)."""
import sys, time
import torch
sys.path.insert(0, ".")
from cca_zoo_b200 import ops

def timeit(fn, warm=2, it=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it

torch.manual_seed(0)
n, d = 100000, 1024
X1 = torch.randn(n, d, device="cuda"); X2 = torch.randn(n, d, device="cuda")
F = n * 2 * d * (2 * d + 1)
for prec in ["tf32", "tf32x3"]:
    t = timeit(lambda: ops.moments([X1, X2], precision=prec))
    print(f"moments {prec} n={n} D={2*d}: {t:.3f} ms  -> {F/t/1e9:.1f} TFLOP/s algorithmic", flush=True)
mom = ops.moments([X1, X2], precision="tf32x3")
t = timeit(lambda: ops.covariance(mom, [d, d], n, dtype=torch.float32))
print(f"covariance finalize: {t:.3f} ms")
Cm, _ = ops.covariance(mom, [d, d], n, dtype=torch.float32)
A = torch.stack([Cm[:d, :d], Cm[d:, d:]]).contiguous()
for dt in [torch.float32, torch.float64]:
    Ad = A.to(dt)
    t0 = time.time(); ev, Vt, info = ops.syevj(Ad, return_info=True); torch.cuda.synchronize(); t1 = time.time()
    t = timeit(lambda: ops.syevj(Ad), warm=0, it=2)
    V = Vt[0].double(); resid = (Ad[0].double() @ V.T - V.T * ev[0].double()).abs().max().item()
    orth = (V @ V.T - torch.eye(d, device="cuda", dtype=torch.float64)).abs().max().item()
    print(f"syevj {dt} batch=2 n={d}: {t:.2f} ms, sweeps={info['sweeps']} offdiag={info['offdiag']:.2e} resid={resid:.2e} orth={orth:.2e}", flush=True)
t = timeit(lambda: ops.gemm(A[0], A[1]))
print(f"gemm fp32 1024^3: {t:.3f} ms -> {2*1024**3/t/1e9:.2f} TFLOP/s")
for nn in [64, 256, 512, 2048]:
    g = torch.randn(2 * nn, nn, device="cuda", dtype=torch.float32)
    S = (g.T @ g / (2 * nn))
    ev, Vt, info = ops.syevj(S, return_info=True)
    t = timeit(lambda: ops.syevj(S), warm=0, it=2)
    print(f"syevj f32 n={nn}: {t:.2f} ms sweeps={info['sweeps']}", flush=True)
