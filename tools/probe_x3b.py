"""3xTF32: accumulator-run length (number of K splits) x operand split -> kernel time and parity at config 2."""
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from cca_zoo_b200 import ops, _lib
from cca_zoo_b200.linear import rCCA
from oracle import restatement as R
import bench
lib = _lib.load()
n, d, k = 100000, 1024, 64
views = bench.make_views(1000)
X = np.hstack(views).astype(np.float64); mu = X.mean(0); X -= mu
C64 = X.T @ X / (n - 1); del X
w_ref, sv = R.cov_rcca_fit(C64, [d, d], k, 0.1, n)
dev = [torch.from_numpy(v).cuda() for v in views]
lib.ccab_profile_moments(1)
for split in [0, 1]:
    for fs in [13, 25, 49]:
        ops.debug_set("x3_split", split); ops.debug_set("force_splits", fs)
        for _ in range(3): ops.moments(dev, "tf32x3")
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5): ops.moments(dev, "tf32x3")
        torch.cuda.synchronize(); call = (time.perf_counter() - t0) / 5 * 1e3
        kms = lib.ccab_profile_moments_last_ms()
        est = rCCA(latent_dimensions=k, c=0.1).fit(dev)
        w = [x.astype(np.float64) for x in est.weights_]
        ws = R.align_signs(w, w_ref)
        pv = np.concatenate([np.linalg.norm(a - b, axis=0) / np.linalg.norm(b, axis=0) for a, b in zip(ws, w_ref)])
        print(f"split={'rn' if split else 'trunc'} S={fs}: moments call {call:.3f} ms (tcgen05 kernel {kms:.3f}) | weights max {pv.max():.2e} median {np.median(pv):.2e}", flush=True)
