"""BASELINE.json full sizes (config 2: n=100000, d=[1024,1024], k=64, float32): size-independent properties.

The oracle cannot run these sizes in test time, so the CUDA path is checked through identities that hold for
any input: additivity of the moments over row shards, quadratic scaling, the trace identity, exact symmetry,
and -- for the fitted model -- W_i^T R_i W_i = I and W_1^T C_12 W_2 = diag(sigma), sigma descending in [0, 1/(1-c)].
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N, D1, K, C_RIDGE = 100_000, 1024, 64, 0.1


@pytest.fixture(scope="module")
def views():
    g = torch.Generator(device="cuda").manual_seed(0)
    z = torch.randn(N, K, generator=g, device="cuda")
    out = []
    for _ in range(2):
        w = torch.randn(K, D1, generator=g, device="cuda")
        out.append(z @ w + 22.6 * torch.randn(N, D1, generator=g, device="cuda") + 0.5)
    return out


@pytest.mark.parametrize("precision,tol", [("tf32x3", 2e-5), ("tf32", 2e-3)])
def test_moment_identities_at_full_size(views, precision, tol):
    from cca_zoo_b200 import ops

    dims = [D1, D1]
    whole = ops.moments(views, precision=precision)
    Cw, mean_w = ops.covariance(whole, dims, N, dtype=torch.float64)
    assert torch.equal(Cw, Cw.T)
    # additivity over row shards (the multi-GPU contract) at three uneven cut points
    cuts = [0, 33_333, 70_001, N]
    parts = sum(ops.moments([v[a:b] for v in views], precision=precision) for a, b in zip(cuts[:-1], cuts[1:]))
    Cp, mean_p = ops.covariance(parts, dims, N, dtype=torch.float64)
    scale = Cw.diagonal().max()
    assert (Cw - Cp).abs().max() < tol * scale
    assert (mean_w - mean_p).abs().max() < tol
    # trace identity: sum of per-column variances (float64 reduction on the device as the checker)
    var = torch.cat([v.double().var(dim=0, unbiased=True) for v in views])
    assert ((Cw.diagonal() - var).abs() / var).max() < 50 * tol
    # quadratic scaling: cov(2 X) = 4 cov(X) (exact in binary floating point up to the split arithmetic)
    C2, _ = ops.covariance(ops.moments([2.0 * v for v in views], precision=precision), dims, N, dtype=torch.float64)
    assert (C2 - 4.0 * Cw).abs().max() < 1e-6 * scale


def test_fit_identities_at_full_size(views):
    from cca_zoo_b200 import ops
    from cca_zoo_b200.linear import rCCA

    est = rCCA(latent_dimensions=K, c=C_RIDGE).fit(views)
    assert est.n_samples_ == N and est.weights_[0].shape == (D1, K) and est.weights_[0].dtype == np.float32
    Cm, _ = ops.covariance(ops.moments(views, precision="tf32x3"), [D1, D1], N, dtype=torch.float64)
    Cm = Cm.cpu().numpy()
    W1, W2 = (w.astype(np.float64) for w in est.weights_)
    eye = np.eye(K)
    for W, blk in ((W1, Cm[:D1, :D1]), (W2, Cm[D1:, D1:])):
        R = (1 - C_RIDGE) * blk + C_RIDGE * np.eye(D1)
        assert np.abs(W.T @ R @ W - eye).max() < 2e-3          # the rCCA constraint (_rcca.py:24-27)
    S = W1.T @ Cm[:D1, D1:] @ W2
    sig = np.abs(np.diag(S))
    assert np.abs(S - np.diag(np.diag(S))).max() < 2e-3          # cross-covariance of the variates is diagonal
    assert np.all(sig[:-1] >= sig[1:] - 1e-4) and sig.max() <= 1.0 / (1.0 - C_RIDGE) + 1e-3
    # means (the column sums ride on the tensor-core pipeline as an extra MMA)
    ref_mean = views[0].double().mean(dim=0).cpu().numpy()
    np.testing.assert_allclose(est.means_[0], ref_mean, rtol=1e-5, atol=1e-5)
