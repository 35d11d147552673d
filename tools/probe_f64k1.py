"""float64 moment kernel (DMMA) on a slice of the config-5 shard: time and algorithmic TFLOP/s."""
import sys, torch
sys.path.insert(0, ".")
from cca_zoo_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 62500
dims = [2048, 2048]
g = torch.Generator(device="cuda").manual_seed(1)
v = [torch.randn(n, d, device="cuda", dtype=torch.float64, generator=g) for d in dims]
for _ in range(2): mom = ops.moments(v, "exact")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(5): mom = ops.moments(v, "exact")
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
D = sum(dims)
F = n * D * (D + 1)
X = torch.cat([t[:4096] for t in v], 1)
m4 = ops.moments([t[:4096] for t in v], "exact")
M = m4[:D * D].view(D, D)
ref = X.T @ X
err = ((torch.triu(M) - torch.triu(ref)).abs().max() / ref.abs().max()).item()
print(f"f64 moments n={n} D={D}: {ms:.2f} ms -> {F / ms / 1e9:.2f} TFLOP/s algorithmic (symmetric), max rel err (4096-row check) {err:.1e}")
