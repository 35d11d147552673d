import sys, time
import torch
sys.path.insert(0, ".")
from cca_zoo_b200 import ops
torch.manual_seed(0)
def run(n, dt, inner, batch=1, unfused=0):
    ops.debug_set("jacobi_inner_sweeps", inner)
    ops.debug_set("jacobi_force_unfused", unfused)
    g = torch.randn(batch, 3 * n, n, device="cuda", dtype=torch.float64)
    A = (g.transpose(1, 2) @ g / (3 * n)).to(dt)
    ev, Vt, info = ops.syevj(A, return_info=True)
    torch.cuda.synchronize(); t0 = time.time()
    ev, Vt, info = ops.syevj(A, return_info=True)
    torch.cuda.synchronize(); t = (time.time() - t0) * 1e3
    V = Vt[0].double(); A0 = A[0].double()
    resid = (A0 @ V.T - V.T * ev[0].double()).abs().max().item()
    orth = (V @ V.T - torch.eye(n, device="cuda", dtype=torch.float64)).abs().max().item()
    ref = torch.linalg.eigvalsh(A0).flip(0)
    everr = ((ev[0].double() - ref).abs().max() / ref.abs().max()).item()
    print(f"n={n} {str(dt)[6:]} batch={batch} inner={inner} unfused={unfused}: {t:8.2f} ms sweeps={info['sweeps']:2d} offdiag={info['offdiag']:.1e} resid={resid:.1e} orth={orth:.1e} evrel={everr:.1e}", flush=True)
for unf in [1, 0]:
    for inner in [1, 2]:
        run(1024, torch.float32, inner, batch=2, unfused=unf)
run(1024, torch.float64, 1, batch=2)
run(1024, torch.float32, 1, batch=1)
run(256, torch.float32, 1); run(2048, torch.float32, 1); run(2048, torch.float64, 1)
run(64, torch.float32, 1); run(64, torch.float64, 1); run(512, torch.float32, 1, batch=4)
