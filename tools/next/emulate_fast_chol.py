"""Lane-level emulation (numpy) of the warp-synchronous 32x32 Cholesky / triangular inverse and of the block
algorithms planned for the solver stage -- validates the data flow (which lane owns what, the
shuffle sources, the order of updates) before it is written as CUDA in tools/next/fast_chol.cu.

    python tools/next/emulate_fast_chol.py
"""
import numpy as np

W = 32


def shfl(vals, src):
    """__shfl_sync: every lane reads lane `src`'s value (vals is the per-lane array of one register)."""
    return np.full(W, vals[src])


def warp_potrf32(A):
    """Lane i owns row i of the lower triangle in registers a[i, 0..31].  Returns L (rows per lane)."""
    a = np.tril(A).copy()                     # a[i, j]: register j of lane i
    lanes = np.arange(W)
    for k in range(W):
        piv = shfl(a[:, k], k)
        d = np.sqrt(piv)
        dinv = 1.0 / d
        lik = np.where(lanes > k, a[:, k] * dinv, np.where(lanes == k, d, 0.0))
        a[:, k] = np.where(lanes >= k, lik, a[:, k])
        for j in range(k + 1, W):
            ljk = shfl(lik, j)
            a[:, j] = np.where(lanes >= j, a[:, j] - lik * ljk, a[:, j])
    return np.tril(a)


def warp_trtri32(L):
    """Lane i owns row i of L (registers l[i, :]) and ends up owning row i of X = L^-1."""
    l = L.copy()
    lanes = np.arange(W)
    x = np.eye(W)
    for k in range(W):
        inv = 1.0 / l[k, k]
        for j in range(k + 1):
            x[k, j] = x[k, j] * inv                        # lane k only
        for j in range(k + 1):
            xkj = shfl(x[:, j], k)
            x[:, j] = np.where(lanes > k, x[:, j] - l[:, k] * xkj, x[:, j])
    return x


def block64_factor_and_inverse(A, nb=64):
    """The 64 x 64 diagonal kernel: 2 x 2 blocks of 32; padded with the identity when nb < 64."""
    S = np.eye(64)
    S[:nb, :nb] = np.tril(A[:nb, :nb]) + np.tril(A[:nb, :nb], -1).T
    X = np.zeros((64, 64))
    L11 = warp_potrf32(S[:32, :32])
    X11 = warp_trtri32(L11)
    L21 = S[32:, :32] @ X11.T                               # step c
    A22 = S[32:, 32:] - L21 @ L21.T                         # step d
    L22 = warp_potrf32(A22)
    X22 = warp_trtri32(L22)
    M = L21 @ X11                                           # step f
    X21 = -X22 @ M
    L = np.zeros((64, 64))
    L[:32, :32], L[32:, :32], L[32:, 32:] = L11, L21, L22
    X[:32, :32], X[32:, :32], X[32:, 32:] = X11, X21, X22
    return L[:nb, :nb], X[:nb, :nb]


def potrf_with_dinv(A, NB=64):
    """Right-looking blocked Cholesky where both triangular solves are GEMMs with the kept inverses."""
    n = A.shape[0]
    A = A.copy()
    dinv = []
    for j0 in range(0, n, NB):
        nb = min(NB, n - j0)
        L, X = block64_factor_and_inverse(A[j0:j0 + nb, j0:j0 + nb], nb)
        A[j0:j0 + nb, j0:j0 + nb] = L
        dinv.append(X)
        if j0 + nb < n:
            P = A[j0 + nb:, j0:j0 + nb] @ X.T               # panel: A_panel L_jj^-T
            A[j0 + nb:, j0:j0 + nb] = P
            A[j0 + nb:, j0 + nb:] -= P @ P.T
    return np.tril(A), dinv


def trtri_doubling(L, dinv, NB=64):
    """X = L^-1 by recursive doubling: diagonal block inverses, then X21 = -X_B (C X_A) per pair and level."""
    n = L.shape[0]
    X = np.zeros_like(L)
    for b, j0 in enumerate(range(0, n, NB)):
        nb = min(NB, n - j0)
        X[j0:j0 + nb, j0:j0 + nb] = dinv[b]
    b = NB
    while b < n:
        for r0 in range(0, n, 2 * b):
            a0, a1 = r0, min(r0 + b, n)
            b0, b1 = a1, min(r0 + 2 * b, n)
            if b0 >= b1:
                continue
            tmp = L[b0:b1, a0:a1] @ X[a0:a1, a0:a1]
            X[b0:b1, a0:a1] = -X[b0:b1, b0:b1] @ tmp
        b *= 2
    return X


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    B = rng.standard_normal((200, 32))
    A = B.T @ B / 200 + 0.1 * np.eye(32)
    L = warp_potrf32(A)
    assert np.allclose(L, np.linalg.cholesky(A), atol=1e-12)
    assert np.allclose(warp_trtri32(L) @ L, np.eye(32), atol=1e-12)
    for nb in (64, 40, 32, 7):
        B = rng.standard_normal((300, nb))
        A = B.T @ B / 300 + 0.05 * np.eye(nb)
        L, X = block64_factor_and_inverse(A, nb)
        assert np.allclose(L, np.linalg.cholesky(A), atol=1e-12), nb
        assert np.allclose(X @ L, np.eye(nb), atol=1e-11), nb
    for n in (64, 96, 200, 1000):
        B = rng.standard_normal((3 * n, n))
        A = B.T @ B / (3 * n) + 0.1 * np.eye(n)
        L, dinv = potrf_with_dinv(A)
        assert np.allclose(L, np.linalg.cholesky(A), atol=1e-10), n
        X = trtri_doubling(L, dinv)
        assert np.allclose(X @ L, np.eye(n), atol=1e-9), n
    print("emulation OK")
