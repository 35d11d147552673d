"""Probe the tcgen05 moment kernel on integer-valued data (exact in TF32): prints error patterns.
Usage: python tools/probe_tf32.py [lbo sbo]   (run each variant in its own process: a trap kills the context)
"""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from cca_zoo_b200 import ops

if len(sys.argv) >= 3:
    ops.debug_set("lbo_bytes", int(sys.argv[1])); ops.debug_set("sbo_bytes", int(sys.argv[2]))
if len(sys.argv) >= 4:
    ops.debug_set("tma_dtype", int(sys.argv[3]))
torch.manual_seed(0)
for (n, dims, prec) in [(64, [128], "tf32"), (256, [128], "tf32"), (256, [256], "tf32"), (1000, [384, 200], "tf32"),
                        (256, [128], "tf32x3"), (1000, [384, 200], "tf32x3")]:
    views = [torch.randint(-4, 5, (n, d), device="cuda").float() for d in dims]
    mom = ops.moments(views, precision=prec)
    torch.cuda.synchronize()
    D = sum(dims)
    X = torch.cat(views, 1).double()
    M = X.T @ X
    s = X.sum(0)
    Dp = int(round((-1 + (1 + 4 * mom.numel()) ** 0.5) / 2))
    Cm, mean = ops.covariance(mom, dims, n, center=False)
    got = Cm * (n - 1)
    err = (got - M).abs()
    _, mean_c = ops.covariance(mom, dims, n, center=True)
    serr = (mean_c * n - s).abs().max().item()
    print(f"n={n} dims={dims} {prec}: max|M err|={err.max().item():.4g} (|M|max={M.abs().max().item():.4g}) "
          f"colsum err={serr:.4g} bad entries={(err > 0.5).sum().item()}/{D*D}", flush=True)
    if err.max() > 0.5:
        bad = (err > 0.5).nonzero()
        print("  first bad:", bad[:6].tolist(), " rows hist(128):", torch.bincount(bad[:, 0] // 32, minlength=D // 32 + 1).tolist()[:16])
