"""Float32 emulation (numpy) of _solvers.topk_svd on the whitened cross-covariance T of bench config 2:
how many products with T^T T, how much oversampling and how many CholQR passes does the residual test
(resid <= 200 eps sigma_1 sqrt(k)) really need?  Guides the round-2 trimming of the solver stage.

    python tools/next/emulate_topk.py            # ~1 min on 8 cores
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402

f32 = np.float32
K = 64


def whitened_T():
    views = bench.make_views(1000)
    X = np.hstack(views).astype(np.float64)
    n = X.shape[0]
    X -= X.mean(axis=0)
    C = (X.T @ X / (n - 1)).astype(f32)
    d = 1024
    Linv = []
    for s in (slice(0, d), slice(d, 2 * d)):
        R = (f32(0.9) * C[s, s] + f32(0.1) * np.eye(d, dtype=f32)).astype(np.float64)
        L = np.linalg.cholesky(R)
        Linv.append(np.linalg.inv(L).astype(f32))
    return (Linv[0] @ C[:d, d:] @ Linv[1].T).astype(f32)


def cholqr(Y, passes):
    for _ in range(passes):
        G = (Y.T @ Y).astype(f32)
        L = np.linalg.cholesky(G.astype(np.float64)).astype(f32)
        Y = np.linalg.solve(L.astype(np.float64), Y.T.astype(np.float64)).T.astype(f32)
    return Y


def run(T, oversample, iters, ortho_every, seed=1234):
    d1, d2 = T.shape
    p = K + oversample
    rng = np.random.default_rng(seed)
    Z = rng.standard_normal((d2, p)).astype(f32)
    n_orth = 0
    for it in range(iters):
        Y = T @ Z
        Z = T.T @ Y
        last = it == iters - 1
        if last or (it + 1) % ortho_every == 0:
            Z = cholqr(Z, 2 if last else 1)
            n_orth += 2 if last else 1
        else:
            Z = Z / np.linalg.norm(Z, axis=0)          # column scaling only (no Gram / Cholesky / solve)
    Y = T @ Z
    sig2, Vy = np.linalg.eigh((Y.T @ Y).astype(np.float64))
    sig2, Vy = sig2[::-1], Vy[:, ::-1].astype(f32)
    sig = np.sqrt(np.maximum(sig2, 0)).astype(f32)
    U = (Y @ Vy[:, :K]) / sig[:K]
    V = Z @ Vy[:, :K]
    E = U.T @ T - (V * sig[:K]).T
    resid = np.linalg.norm(E)
    limit = 200 * np.finfo(f32).eps * sig[0] * np.sqrt(K)
    return resid, limit, n_orth, sig


if __name__ == "__main__":
    t0 = time.time()
    T = whitened_T()
    sv = np.linalg.svd(T.astype(np.float64), compute_uv=False)
    print(f"T built in {time.time() - t0:.0f} s; sigma_1 {sv[0]:.4f} sigma_64 {sv[63]:.4f} sigma_65 {sv[64]:.4f} "
          f"sigma_96 {sv[95]:.4f} sigma_128 {sv[127]:.4f}")
    print("oversample iters ortho_every | resid / limit | CholQR passes | max rel err of sigma_1..64")
    for oversample in (16, 32, 64):
        for iters in (2, 3, 4, 5):
            for every in (1, 2, 5):
                if every > iters:
                    continue
                try:
                    resid, limit, n_orth, sig = run(T, oversample, iters, every)
                except np.linalg.LinAlgError:
                    print(f"{oversample:9d} {iters:5d} {every:11d} | Gram matrix of the block lost positive definiteness")
                    continue
                err = np.max(np.abs(sig[:K] - sv[:K]) / sv[:K])
                print(f"{oversample:9d} {iters:5d} {every:11d} | {resid / limit:8.3f} {'ok ' if resid <= limit else 'NO '}"
                      f"| {n_orth:2d} | {err:.1e}")
