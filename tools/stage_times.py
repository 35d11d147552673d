import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from cca_zoo_b200 import ops, _solvers
from cca_zoo_b200.datasets import joint_data
torch.manual_seed(0)
n, d, k = 100000, 1024, 64
views = [torch.from_numpy(v).cuda() for v in joint_data(2, n, k, [d, d], 2.0 / 1024, 0, np.float32)]
class T:
    def __init__(s, name): s.name = name
    def __enter__(s):
        torch.cuda.synchronize(); s.t = time.perf_counter()
    def __exit__(s, *a):
        torch.cuda.synchronize(); print(f"{s.name:34s} {(time.perf_counter()-s.t)*1e3:8.3f} ms")
for rep in range(2):
    print("--- rep", rep)
    with T("moments tf32x3"): mom = ops.moments(views, "tf32x3")
    with T("covariance"): C, mean = ops.covariance(mom, [d, d], n, True, torch.float32)
    s1, s2 = slice(0, d), slice(d, 2 * d)
    Ls = []
    with T("R build + potrf x2 (+2 syncs)"):
        for s in (s1, s2):
            R = 0.9 * C[s, s]; R.diagonal().add_(0.1)
            dmax = float(R.diagonal().max().item())
            info = ops.potrf_(R, pivot_tol=1e5 * 1.2e-7 * dmax); assert int(info.item()) == 0
            Ls.append(R)
    with T("T = L1^-1 C12 L2^-T"):
        Tm = C[s1, s2].contiguous(); ops.trsm_(Ls[0], Tm, "left"); ops.trsm_(Ls[1], Tm, "right", True)
    with T("topk_svd total"): res = _solvers.topk_svd(Tm, k)
    Z = torch.randn(d, 128, device="cuda")
    with T("  one cholqr (1024x128)"): _solvers._cholqr_(Z)
    with T("  one gemm T Z"): Y = ops.gemm(Tm, Z)
    Yt = ops.gemm(Z, Tm, transa=True, transb=True)
    with T("  gesvj (128 x 1024)"): ops.gesvj(Yt)
    sig, Ut, Vt = res
    with T("back-substitution x2"):
        w1 = Ut.T.contiguous(); w2 = Vt.T.contiguous()
        ops.trsm_(Ls[0], w1, "left", True); ops.trsm_(Ls[1], w2, "left", True)
    with T("D2H weights"): a = w1.cpu().numpy(); b = w2.cpu().numpy()
