"""Entry-wise accuracy of the 3xTF32 covariance under the two operand splits, config-2 size."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from cca_zoo_b200 import ops
from cca_zoo_b200.datasets import joint_data
n, d = 100000, 1024
views = joint_data(2, n, 64, [d, d], 2.0 / 1024, 0, np.float32)
X = np.hstack(views).astype(np.float64); mu = X.mean(0); Xc = X - mu
C64 = Xc.T @ Xc / (n - 1)
sc = np.sqrt(np.outer(np.diag(C64), np.diag(C64)))
dev = [torch.from_numpy(v).cuda() for v in views]
for name, split in [("residual(trunc)", 0), ("rn hi/lo", 1)]:
    ops.debug_set("x3_split", split)
    for fs in [0, 2]:
        ops.debug_set("force_splits", fs)
        C, _ = ops.covariance(ops.moments(dev, "tf32x3"), [d, d], n, dtype=torch.float64)
        E = (C.cpu().numpy() - C64) / sc
        dg = np.diag(E)
        off = E[~np.eye(2 * d, dtype=bool)]
        print(f"{name:16s} splits={'auto(13)' if fs == 0 else fs}: diag rel err mean {dg.mean():+.2e} std {dg.std():.2e} | offdiag normalised err mean {off.mean():+.2e} std {off.std():.2e} max {np.abs(off).max():.2e}", flush=True)
ops.debug_set("force_splits", 0)
C, _ = ops.covariance(ops.moments(dev, "exact"), [d, d], n, dtype=torch.float64)
E = (C.cpu().numpy() - C64) / sc; dg = np.diag(E); off = E[~np.eye(2 * d, dtype=bool)]
print(f"exact fp32 FMA: diag mean {dg.mean():+.2e} std {dg.std():.2e} | offdiag std {off.std():.2e} max {np.abs(off).max():.2e}")
