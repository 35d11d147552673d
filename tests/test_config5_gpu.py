"""BASELINE.json configs[4] (GCCA, 8 views x d=2048, k=128, float64, n=5e5 over 8 GPUs): ONE RANK'S ROW SHARD
(n=62500, the weak-scaling unit) at the full width D = 16384.

Neither the reference (its n x n matrix would need 2 TB at n=5e5; 31 GB for this shard alone) nor the numpy
oracle (a 16384 x 16384 generalised eigenproblem) runs this in test time, so the fit is checked through
size-independent properties with cuBLAS / cuSOLVER float64 (torch) as the independent checker:

  * K1 (fp64 DMMA) covariance == X^T X route of torch.matmul, entrywise;
  * GCCA's fixed-point equations in covariance form (derived from cca_zoo/linear/_gcca.py:94-110 with
    a_i = X_i^T T): C W = B W diag(rho), B = blkdiag(C_ii), rho = sigma/(n-1) descending in (1, m];
    W^T B W = diag(rho)/(n-1)  (this is T^T T = I);
  * the k found eigenpairs are the LARGEST ones: power iteration on B^-1 C in the B-orthogonal complement of
    the found vectors stays below rho_k.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N, M, D1, K = 62_500, 8, 2048, 128


def make_shard(n=N, seed=0, device="cuda"):
    """Rows of one population: K shared latents, per-view loadings with per-view R^2 ~ 0.8, unit noise, offsets."""
    g = torch.Generator(device=device).manual_seed(20240924)           # loadings: the same on every rank
    loads = [torch.randn(K, D1, generator=g, device=device, dtype=torch.float64) * (4.0 / D1) ** 0.5
             for _ in range(M)]
    g = torch.Generator(device=device).manual_seed(seed)               # rows: per rank
    z = torch.randn(n, K, generator=g, device=device, dtype=torch.float64)
    views = []
    for a in loads:
        x = torch.randn(n, D1, generator=g, device=device, dtype=torch.float64)
        x.addmm_(z, a).add_(0.25)
        views.append(x)
    return views


def check_gcca_properties(views, weights, k=K, allreduce=None):
    """Returns a dict of the measured residuals (asserted by the test, reported by tools/config5_shard.py).
    ``allreduce`` (tensor -> None, in-place sum over ranks) makes the checker's covariance the global one when
    the views are a row shard."""
    n, m = views[0].shape[0], len(views)
    X = torch.cat(views, dim=1)
    s = X.sum(dim=0)
    G = X.T @ X                                                        # cuBLAS float64: the checker
    del X
    if allreduce is not None:
        cnt = torch.tensor([float(n)], dtype=torch.float64, device=G.device)
        for t in (G, s, cnt):
            allreduce(t)
        n = int(cnt.item())
    C = (G - torch.outer(s, s) / n) / (n - 1)
    del G
    dims = [v.shape[1] for v in views]
    off = np.concatenate([[0], np.cumsum(dims)])
    W = torch.cat([torch.from_numpy(np.ascontiguousarray(w)).to(C.device) for w in weights], dim=0)   # D x k
    BW = torch.cat([C[off[i]:off[i + 1], off[i]:off[i + 1]] @ W[off[i]:off[i + 1]] for i in range(m)], dim=0)
    NB = W.T @ BW                                                      # W^T B W
    rho = (n - 1) * NB.diagonal()
    offdiag = (NB - torch.diag(NB.diagonal())).abs().max() / NB.diagonal().abs().max()
    CW = C @ W
    resid = torch.linalg.norm(CW - BW * rho) / torch.linalg.norm(CW)
    # top-ness: power iteration on B^-1 C, deflated against the found vectors (B-orthonormalised)
    chol = [torch.linalg.cholesky(C[off[i]:off[i + 1], off[i]:off[i + 1]]) for i in range(m)]
    Wn = W / NB.diagonal().sqrt()
    BWn = BW / NB.diagonal().sqrt()
    g = torch.Generator(device=C.device).manual_seed(7)
    v = torch.randn(C.shape[0], 4, generator=g, device=C.device, dtype=C.dtype)
    best = 0.0
    for _ in range(40):
        v = v - Wn @ (BWn.T @ v)
        cv = C @ v
        bv = torch.cat([C[off[i]:off[i + 1], off[i]:off[i + 1]] @ v[off[i]:off[i + 1]] for i in range(m)], dim=0)
        best = float(((v * cv).sum(0) / (v * bv).sum(0)).max())
        v = torch.cat([torch.cholesky_solve(cv[off[i]:off[i + 1]], chol[i]) for i in range(m)], dim=0)
        v = v / torch.linalg.norm(v, dim=0)
    return {"C": C, "rho": rho.cpu().numpy(), "offdiag_rel": float(offdiag), "eig_resid_rel": float(resid),
            "next_rayleigh": best}


def test_gcca_config5_rank_shard():
    free, _ = torch.cuda.mem_get_info()
    if free < 48 << 30:
        pytest.skip("needs ~40 GB of device memory")
    from cca_zoo_b200 import ops
    from cca_zoo_b200.linear import GCCA

    views = make_shard()
    est = GCCA(latent_dimensions=K).fit(views)
    assert len(est.weights_) == M and all(w.shape == (D1, K) and w.dtype == np.float64 for w in est.weights_)
    assert est.n_samples_ == N
    res = check_gcca_properties(views, est.weights_)
    # K1 on the fp64 tensor pipe against cuBLAS at full width
    Cours, _ = ops.covariance(ops.moments(views), [D1] * M, N, dtype=torch.float64)
    scale = float(res["C"].diagonal().max())
    assert float((Cours - res["C"]).abs().max()) < 1e-11 * scale
    del Cours
    rho = res["rho"]
    assert np.all(np.diff(rho) <= 1e-9 * rho[0]) and rho[-1] > 1.0 and rho[0] <= M * (1 + 1e-9)
    assert res["offdiag_rel"] < 1e-8, res
    assert res["eig_resid_rel"] < 1e-8, res
    assert res["next_rayleigh"] < rho[-1], res            # nothing larger was left behind
    assert rho[-1] > 3.0                                   # the planted shared subspace (rho ~ 1 + 7 * 0.8)
