"""tf32x3b moment kernel: persistent (overlapped epilogue) vs one unit per CTA pair -- time, agreement, parity at config 2."""
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from cca_zoo_b200 import ops, _lib
from cca_zoo_b200.linear import rCCA
from oracle import restatement as R
import bench
lib = _lib.load()
n, d, k = 100000, 1024, 64
views = bench.make_views(1000)
X = np.hstack(views).astype(np.float64)
s64 = X.sum(0); M64 = None
mu = X.mean(0); X -= mu
C64 = X.T @ X / (n - 1); del X
w_ref, sv = R.cov_rcca_fit(C64, [d, d], k, 0.1, n)
dev = [torch.from_numpy(v).cuda() for v in views]
lib.ccab_profile_moments(1)
out = {}
for oneshot in [1, 0, 1, 0]:
    ops.debug_set("x3b_oneshot", oneshot)
    for _ in range(3): m = ops.moments(dev, "tf32x3b")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): m = ops.moments(dev, "tf32x3b")
    torch.cuda.synchronize(); call = (time.perf_counter() - t0) / 10 * 1e3
    kms = lib.ccab_profile_moments_last_ms()
    out[oneshot] = m.clone()
    D = 2 * d
    s = m[D * D:D * D + D].cpu().numpy()
    est = rCCA(latent_dimensions=k, c=0.1).fit(dev)
    w = [x.astype(np.float64) for x in est.weights_]
    ws = R.align_signs(w, w_ref)
    pv = np.concatenate([np.linalg.norm(a - b, axis=0) / np.linalg.norm(b, axis=0) for a, b in zip(ws, w_ref)])
    print(f"oneshot={oneshot}: moments call {call:.3f} ms (tcgen05 kernel {kms:.3f}) | sums max abs err {np.abs(s - s64).max():.3e} "
          f"| weights max {pv.max():.2e} median {np.median(pv):.2e} route {est._fit_info.get('route')}", flush=True)
a, b = out[0], out[1]
D = 2 * d
Ma, Mb = a[:D * D].view(D, D), b[:D * D].view(D, D)
print("M persistent vs oneshot: max abs diff", (Ma - Mb).abs().max().item(), "bitwise equal:", bool(torch.equal(Ma, Mb)))
