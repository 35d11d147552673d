"""Config-2 rCCA fit from device-resident views: block width (oversampling) x first-try iteration count -> time, parity."""
import os, sys, numpy as np, torch
sys.path.insert(0, ".")
from cca_zoo_b200.linear import rCCA
from cca_zoo_b200 import ops
from oracle import restatement as R
import bench
n, d, k = 100000, 1024, 64
views = bench.make_views(1000)
X = np.hstack(views).astype(np.float64); X -= X.mean(0)
C64 = X.T @ X / (n - 1); del X
w_ref, sv = R.cov_rcca_fit(C64, [d, d], k, 0.1, n)
dev = [torch.from_numpy(v).cuda() for v in views]
combos = [(32, 5), (16, 5), (16, 3), (24, 4), (32, 3)] if len(sys.argv) < 2 else [tuple(map(int, a.split(","))) for a in sys.argv[1:]]
for combo in combos:
    over, it = combo[0], combo[1]
    split = combo[2] if len(combo) > 2 else 1
    ops.debug_set("gemm_split", split)
    os.environ["CCAB_FIT_OVERSAMPLE"] = str(over); os.environ["CCAB_FIT_ITERS"] = str(it)
    for _ in range(3): est = rCCA(latent_dimensions=k, c=0.1).fit(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): est = rCCA(latent_dimensions=k, c=0.1).fit(dev)
    e1.record(); torch.cuda.synchronize()
    w = [x.astype(np.float64) for x in est.weights_]
    ws = R.align_signs(w, w_ref)
    pv = np.concatenate([np.linalg.norm(a - b, axis=0) / np.linalg.norm(b, axis=0) for a, b in zip(ws, w_ref)])
    print(f"oversample={over} iters={it} split={split}: {e0.elapsed_time(e1) / 10:.3f} ms/fit | weights max {pv.max():.2e} median {np.median(pv):.2e} | {est._fit_info}", flush=True)
