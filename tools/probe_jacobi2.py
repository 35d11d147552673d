import sys, time
import torch
sys.path.insert(0, ".")
from cca_zoo_b200 import ops
torch.manual_seed(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dt = torch.float64 if (len(sys.argv) > 2 and sys.argv[2] == "f64") else torch.float32
unf = int(sys.argv[3]) if len(sys.argv) > 3 else 0
ops.debug_set("jacobi_inner_sweeps", 1)
ops.debug_set("jacobi_force_unfused", unf)
g = torch.randn(2, 3 * n, n, device="cuda", dtype=torch.float64)
A = (g.transpose(1, 2) @ g / (3 * n)).to(dt)
ev, Vt, info = ops.syevj(A, return_info=True)
torch.cuda.synchronize()
print(info)
