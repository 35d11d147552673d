"""One eigen-route fit (per-view one-sided block Jacobi + Jacobi SVD of the whitened cross-covariance): the profiling
target for jacobi_round_fused_kernel."""
import sys
import numpy as np
sys.path.insert(0, ".")
from cca_zoo_b200.linear import rCCA
rng = np.random.default_rng(0)
z = rng.standard_normal((20000, 8))
views = [(z @ rng.standard_normal((8, d)) * 0.5 + rng.standard_normal((20000, d))).astype(np.float32) for d in (512, 512)]
for _ in range(2):
    est = rCCA(latent_dimensions=8, c=0.1, solver="eigen").fit(views)
print("eigen route ok", est.weights_[0].shape)
