import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from cca_zoo_b200.linear import MCCA, GCCA
rng_w = np.random.default_rng(1)
n, m, d, k = 125000, 4, 512, 32
W = [rng_w.standard_normal((d, k)) for _ in range(m)]
rng = np.random.default_rng(2)
z = rng.standard_normal((n, k))
views = [torch.from_numpy((z @ w.T + rng.standard_normal((n, d)) * np.sqrt(256.0)).astype(np.float32)).cuda() for w in W]
for cls, kw in [(MCCA, {}), (GCCA, {})]:
    for solver in ["cholesky", "eigen"]:
        est = cls(latent_dimensions=k, solver=solver, **kw)
        est.fit(views); torch.cuda.synchronize()
        t0 = time.perf_counter(); est.fit(views); torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
        sc = est.score([v[:20000].cpu().numpy() for v in views])
        print(f"{cls.__name__} {solver}: fit {dt:.1f} ms, score[:3]={sc[:3]}", flush=True)
