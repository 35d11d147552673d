import os, sys, time, numpy as np, torch
sys.path.insert(0, ".")
import bench
from cca_zoo_b200.linear import rCCA
from oracle import restatement as R
views = bench.make_views(1000)
dev = [torch.from_numpy(v).cuda() for v in views]
ref = None
for over, iters in [(64, 5), (32, 5), (32, 4), (32, 3), (64, 4), (64, 3), (48, 4)]:
    os.environ["CCAB_TOPK_OVERSAMPLE"] = str(over); os.environ["CCAB_TOPK_ITERS"] = str(iters); os.environ["CCAB_DEBUG_TOPK"] = "1"
    est = rCCA(latent_dimensions=64, c=0.1); est.fit(dev)
    os.environ.pop("CCAB_DEBUG_TOPK")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): est.fit(dev)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5 * 1e3
    w = [x.astype(np.float64) for x in est.weights_]
    if ref is None: ref = w
    print(f"oversample={over} iters={iters}: fit {dt:.2f} ms, weights vs first config {R.max_rel_err_per_vector(w, ref):.2e}", flush=True)
