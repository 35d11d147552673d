"""Timings of the round-2 solver building blocks (CUDA events, warm, median of 20)."""
import sys
import torch
sys.path.insert(0, ".")
from cca_zoo_b200 import ops


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


g = torch.Generator(device="cuda").manual_seed(0)
for (m, n, k) in [(1024, 1024, 1024), (1024, 96, 1024), (96, 96, 1024), (4096, 512, 512), (2048, 2048, 2048), (896, 128, 128)]:
    A = torch.randn(m, k, generator=g, device="cuda"); B = torch.randn(k, n, generator=g, device="cuda")
    out = torch.empty(m, n, device="cuda")
    t_tc = timeit(lambda: ops.gemm_tc(A, B, out=out))
    t_fma = timeit(lambda: ops.gemm(A, B, out=out))
    t_cublas = timeit(lambda: torch.matmul(A, B, out=out))
    print(f"gemm {m}x{n}x{k}: tc {t_tc:8.1f} us ({2*m*n*k/t_tc/1e6:7.1f} TF/s)  fma {t_fma:8.1f} us  cublas-fp32 {t_cublas:8.1f} us")

for dtype in (torch.float32, torch.float64):
    for n, batch in [(512, 2), (1024, 2), (512, 4), (2048, 1)]:
        X = torch.randn(batch, 2 * n, n, generator=g, device="cuda", dtype=dtype)
        A0 = X.transpose(1, 2) @ X / (2 * n) + 0.1 * torch.eye(n, device="cuda", dtype=dtype)
        def new():
            A = A0.clone(); ops.potrf_inv_(A)
        def old():
            for b in range(batch):
                A = A0[b].clone(); ops.potrf_(A); E = torch.eye(n, device="cuda", dtype=dtype); ops.trsm_(A, E, side="left")
        print(f"chol+inv {dtype} n={n} batch={batch}: new {timeit(new, 10):9.1f} us   old {timeit(old, 5):9.1f} us")

for dtype, n in [(torch.float32, 64), (torch.float32, 96), (torch.float32, 128), (torch.float64, 64), (torch.float64, 96)]:
    Y = torch.randn(1024, n, generator=g, device="cuda", dtype=dtype)
    H = Y.T @ Y
    t_new = timeit(lambda: ops.syevj_small(H), 10)
    t_old = timeit(lambda: ops.syevj(H), 5)
    _, _, info = ops.syevj_small(H)
    print(f"syev {dtype} n={n}: small {t_new:9.1f} us (sweeps {int(info[0])})   block-jacobi {t_old:9.1f} us")
