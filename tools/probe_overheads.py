import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from cca_zoo_b200 import ops, _solvers, _lib
lib = _lib.load()
torch.manual_seed(0)
n, d, k = 100000, 1024, 64
views = [torch.randn(n, d, device="cuda") for _ in range(2)]
def wall(fn, it=8):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(it):
        torch.cuda.synchronize(); t = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t) * 1e3)
    return min(ts), sum(ts) / len(ts)
lib.ccab_profile_moments(1)
for prec in ["tf32", "tf32x3"]:
    print(prec, "moments wall min/avg ms:", wall(lambda: ops.moments(views, prec)), "kernel", lib.ccab_profile_moments_last_ms())
print("torch.empty 1.7GB:", wall(lambda: torch.empty(1_700_000_000, dtype=torch.uint8, device="cuda")))
print("alloc stats: num cudaMalloc", torch.cuda.memory_stats()["num_device_alloc"], "retries", torch.cuda.memory_stats()["num_alloc_retries"])
mom = ops.moments(views, "tf32x3")
print("covariance:", wall(lambda: ops.covariance(mom, [d, d], n, True, torch.float32)))
C, _ = ops.covariance(mom, [d, d], n, True, torch.float32)
R = (0.9 * C[:d, :d]).contiguous(); R.diagonal().add_(0.1)
print("potrf 1024:", wall(lambda: ops.potrf_(R.clone())))
L = R.clone(); ops.potrf_(L)
B = C[:d, d:].contiguous()
print("trsm left 1024x1024:", wall(lambda: ops.trsm_(L, B.clone(), "left")))
print("trsm right 1024x1024:", wall(lambda: ops.trsm_(L, B.clone(), "right", True)))
W = torch.randn(d, 64, device="cuda")
print("trsm left^T 1024x64:", wall(lambda: ops.trsm_(L, W.clone(), "left", True)))
Z = torch.randn(d, 128, device="cuda")
print("gemm Z^T Z (128x1024x128):", wall(lambda: ops.gemm(Z, Z, transa=True)))
G = ops.gemm(Z, Z, transa=True)
print("potrf 128:", wall(lambda: ops.potrf_(G.clone())))
Lg = G.clone(); ops.potrf_(Lg)
print("trsm right 1024x128:", wall(lambda: ops.trsm_(Lg, Z.clone(), "right", True)))
print("gemm 1024x1024x128:", wall(lambda: ops.gemm(B, Z)))
print("clone 1024x128:", wall(lambda: Z.clone()))
print("item sync:", wall(lambda: float(G[0, 0].item())))
