import sys, time, torch
sys.path.insert(0, ".")
from cca_zoo_b200 import ops
torch.manual_seed(0)
n, d = 20000, 1024
v = [torch.randn(n, d, device="cuda", dtype=torch.float64) for _ in range(2)]
X = torch.cat(v, 1); M = X.T @ X
for simt in [1, 0]:
    ops.debug_set("f64_simt", simt)
    for _ in range(2): ops.moments(v, "exact")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): mom = ops.moments(v, "exact")
    torch.cuda.synchronize(); t = (time.perf_counter() - t0) / 3
    C, _ = ops.covariance(mom, [d, d], n, center=False)
    err = ((C * (n - 1) - M).abs().max() / M.abs().max()).item()
    F = n * 2 * d * (2 * d + 1)
    print(f"f64 moments {'SIMT' if simt else 'DMMA'} n={n} D={2*d}: {t*1e3:.2f} ms -> {F/t/1e12:.2f} TFLOP/s algorithmic, max rel err {err:.1e}")
