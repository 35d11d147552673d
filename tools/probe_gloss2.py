"""Where GCCALoss spends its time at small widths: per-op CUDA-event timings."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cca_zoo_b200 import ops  # noqa: E402
from cca_zoo_b200.deep import GCCALoss  # noqa: E402


def t(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out


for n, widths, dt in [(1024, [16, 16, 16], torch.float32), (4096, [64, 64, 64], torch.float32)]:
    zs = [torch.randn(n, w, device="cuda", dtype=dt) for w in widths]
    D = sum(widths)
    ms, mom = t(lambda: ops.moments(zs, precision="exact"))
    print(f"n={n} D={D}: moments {ms:.3f} ms")
    C, _ = ops.covariance(mom, widths, n, dtype=dt)
    ms, _ = t(lambda: ops.syevj(C[:widths[0], :widths[0]].contiguous()))
    print(f"  syevj view block {widths[0]}: {ms:.3f} ms")
    Wt = torch.zeros((D, D), dtype=dt, device="cuda")
    off = 0
    for d in widths:
        lam, Vt = ops.syevj(C[off:off + d, off:off + d].contiguous())
        Wi, _, _ = ops.whiten_rows(lam, Vt, 0.0, floor_add=1e-5, rank_tol=-1.0, lam_floor=0.0)
        Wt[off:off + d, off:off + d] = Wi
        off += d
    K = ops.gemm(ops.gemm(Wt, C), Wt, transb=True, alpha=float(n - 1))
    K = 0.5 * (K + K.T)
    os.environ["CCAB_JACOBI_VERBOSE"] = "1"
    ops.syevj(K)
    os.environ.pop("CCAB_JACOBI_VERBOSE")
    ms, (ev, _) = t(lambda: ops.syevj(K))
    print(f"  syevj K {D}x{D}: {ms:.3f} ms   evals {ev[0].item():.1f} .. {ev[-1].item():.1f}")
    ms, _ = t(lambda: ops.syevj(K / float(n - 1)))
    print(f"  syevj K/(n-1): {ms:.3f} ms")
    fn = GCCALoss()
    zr = [z.clone().requires_grad_(True) for z in zs]
    ms, loss = t(lambda: fn(zr))
    print(f"  forward {ms:.3f} ms")
    ms, _ = t(lambda: fn(zr).backward())
    print(f"  forward+backward {ms:.3f} ms")
