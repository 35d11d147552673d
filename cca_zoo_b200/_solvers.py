"""Covariance-space solvers on the device (everything after the all-reduce; replicated per rank).

Input: the compact block covariance ``C`` (D x D CUDA tensor) of the hstacked views.  All arithmetic
runs in libccab200 kernels (Jacobi eigensolver / SVD, GEMM, scalings); torch supplies buffers, views
and the handful of scalar read-backs (ranks, floors) that decide shapes on the host.
Algebra: SURVEY.md §3.1-3.3, checked against the reference by oracle/restatement.py (cov_* forms).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def _slices(dims):
    off = np.concatenate([[0], np.cumsum(dims)]).astype(int)
    return [slice(int(off[i]), int(off[i + 1])) for i in range(len(dims))]


def _eps(dtype):
    return float(torch.finfo(dtype).eps)


def _rank_tol(d, dtype):
    """Relative eigenvalue threshold below which a direction of a d x d covariance block is treated as
    numerically null: the covariance-space image of the reference's ``s > 0`` filter
    (cca_zoo/_utils/_linalg.py:30).  d * eps is the noise level of the computed spectrum (entries carry
    O(eps) relative error, the spectral norm of that perturbation grows like d); it does NOT grow with the
    number of samples."""
    return d * _eps(dtype)


def _block_eigh(C, dims):
    """Eigendecomposition of every diagonal block C_ii; equal-sized blocks go in one batched call."""
    sl = _slices(dims)
    lams, vts = [None] * len(dims), [None] * len(dims)
    by_size = {}
    for i, d in enumerate(dims):
        by_size.setdefault(d, []).append(i)
    for d, idx in by_size.items():
        A = torch.stack([C[sl[i], sl[i]] for i in idx]).contiguous()
        ev, evt = ops.syevj(A)
        for b, i in enumerate(idx):
            lams[i], vts[i] = ev[b], evt[b]
    return lams, vts


def _cholqr_(Y, flags, passes=1):
    """Orthonormalise the columns of Y (n x p) in place by CholQR (Gram matrix, Cholesky, triangular solve).
    The device-side Cholesky status flags are appended to ``flags`` (checked once, later, by the caller):
    a non-zero flag means the block lost rank."""
    for _ in range(passes):
        G = ops.gemm(Y, Y, transa=True)
        flags.append(ops.potrf_(G))
        ops.trsm_(G, Y, side="right", trans=True)
    return Y


def topk_svd(T, k, max_rounds=4, iters_per_round=5, oversample=None, seed=1234):
    """Leading k singular triplets of T (d1 x d2) by blocked subspace iteration + Rayleigh-Ritz.

    Z <- orth(T^T orth(T Z)) repeated (one CholQR pass per product: enough to keep the block
    well-conditioned; the last one is done twice); then the Jacobi SVD of the thin block Y = T Z (d1 x p)
    gives U, sigma and V = Z V_y.  Converged when ||T^T U_k - V_k diag(sigma)||_F <= tol sigma_1 sqrt(k) (the
    other residual T v_j - sigma_j u_j vanishes by construction).  Returns (sigma[k], Ut[k,d1], Vt[k,d2]) or
    None if it does not converge (no spectral gap after the block) or a block loses rank: the caller then
    runs the full Jacobi SVD.  One host read-back per round.
    """
    d1, d2 = T.shape
    p = min(min(d1, d2), max(2 * k, k + 32) if oversample is None else k + oversample)
    gen = torch.Generator(device=T.device).manual_seed(seed)
    Z = torch.randn((d2, p), generator=gen, device=T.device, dtype=T.dtype)
    flags = []
    _cholqr_(Z, flags, passes=2)
    tol = 200.0 * _eps(T.dtype)
    for _ in range(max_rounds):
        for it in range(iters_per_round):
            Y = _cholqr_(ops.gemm(T, Z), flags)                   # d1 x p
            Z = _cholqr_(ops.gemm(T, Y, transa=True), flags,      # d2 x p
                         passes=2 if it == iters_per_round - 1 else 1)
        Yt = ops.gemm(Z, T, transa=True, transb=True)       # (T Z)^T : p x d1, rows = columns of Y
        sig, Vy_t, Ut = ops.gesvj(Yt)                       # Y = U diag(sig) Vy^T
        Vt = ops.gemm(Vy_t[:k], Z, transb=True)             # k x d2 : rows of (Z Vy)^T
        E = ops.gemm(Ut[:k], T)                             # rows: u_j^T T
        E -= ops.scale(Vt, rows=sig[:k])
        stats = torch.stack([ops.frobenius_norm(E)[0], sig[0], torch.stack(flags).max().to(T.dtype).reshape(())])
        resid, s1, bad = (float(x) for x in stats.cpu())    # the round's single host read-back
        if bad != 0.0 or not (s1 > 0.0):
            return None
        if resid <= tol * s1 * (k ** 0.5):
            return sig[:k], Ut[:k], Vt
        flags = []
        Z = ops.gemm(Z, Vy_t, transb=True)                  # continue from the Ritz basis (all p vectors)
    return None


def rcca_weights_cholesky(C, dims, n_samples, latent_dimensions, c):
    """rCCA through the Cholesky form of the whitening (same weights as ``rcca_weights`` up to sign):
    R_i = (1-c_i) C_ii + c_i I = L_i L_i^T ; T = L_1^-1 C_12 L_2^-T ; weights = L_i^-T U_k / V_k.
    Returns None when a regularised block is not numerically positive definite (the eigen route, which
    reproduces the reference's rank handling, is used instead)."""
    s1, s2 = _slices(dims)
    Ls = []
    for i, s in enumerate((s1, s2)):
        R = (1.0 - c[i]) * C[s, s]
        R.diagonal().add_(c[i])
        dmax = float(R.diagonal().max().item())
        info = ops.potrf_(R, pivot_tol=_rank_tol(dims[i], C.dtype) * dmax)
        if int(info.item()) != 0:
            return None
        Ls.append(R)
    T = C[s1, s2].contiguous()
    ops.trsm_(Ls[0], T, side="left")                 # L1^-1 C12
    ops.trsm_(Ls[1], T, side="right", trans=True)    # ... L2^-T
    k = min(latent_dimensions, dims[0], dims[1])
    res = topk_svd(T, k) if 4 * k <= min(dims) else None
    if res is None:
        _, Ut, Vt = ops.gesvj(T)
        Ut, Vt = Ut[:k], Vt[:k]
    else:
        _, Ut, Vt = res
    w1 = Ut.T.contiguous()
    w2 = Vt.T.contiguous()
    ops.trsm_(Ls[0], w1, side="left", trans=True)    # L1^-T U_k
    ops.trsm_(Ls[1], w2, side="left", trans=True)
    return [w1, w2]


def rcca_weights(C, dims, n_samples, latent_dimensions, c, solver="auto"):
    """rCCA / CCA / PLS (cca_zoo/linear/_rcca.py:83-101 in covariance form).

    C_ii = V_i L_i V_i^T ; Wt_i = diag(((1-c_i) L_i + c_i)^-1/2) V_i^T (directions with
    lam <= tol*lam_max dropped = the reference's ``s > 0`` filter, _linalg.py:30) ;
    T = Wt_1 C_12 Wt_2^T = U S V^T (one-sided Jacobi) ; weights = Wt_1^T U_k, Wt_2^T V_k.
    """
    if solver == "cholesky" or (solver == "auto" and min(dims) >= 256 and n_samples > max(dims)):
        w = rcca_weights_cholesky(C, dims, n_samples, latent_dimensions, c)
        if w is not None:
            return w
    s1, s2 = _slices(dims)
    lams, vts = _block_eigh(C, dims)
    wts, ranks = [], []
    for i in range(2):
        Wt, _, rank = ops.whiten_rows(lams[i], vts[i], c[i], rank_tol=_rank_tol(dims[i], C.dtype),
                                      max_rank=min(n_samples, dims[i]))
        wts.append(Wt)
        ranks.append(rank)
    r1, r2 = (int(r.item()) for r in ranks)  # host read-back: decides k (_rcca.py:95)
    k = min(latent_dimensions, r1, r2)
    tmp = ops.gemm(wts[0], C[s1, s2])                 # (d1 x d2)
    T = ops.gemm(tmp, wts[1], transb=True)            # (d1 x d2) in whitened coordinates
    # gesvj factors G = T^T given by columns, i.e. the row-major buffer of T itself:
    #   right vectors of G = left singular vectors U of T, left vectors of G = right singular vectors V of T
    _, Ut, Vt = ops.gesvj(T)
    w1 = ops.gemm(wts[0], Ut[:k], transa=True, transb=True)   # (d1 x k)
    w2 = ops.gemm(wts[1], Vt[:k], transa=True, transb=True)   # (d2 x k)
    return [w1, w2]


def mcca_weights(C, dims, latent_dimensions, c, eps):
    """MCCA (cca_zoo/linear/_mcca.py:113-135,141-173): top-k of A v = lam B v, v^T B v = 1 with
    A = (C - blkdiag C_ii)/m, B = blkdiag((1-c_i) C_ii + c_i I)/m (+ eps floor).

    With B_i = V_i diag(b_i) V_i^T and Wt_i = diag(b_i^-1/2) V_i^T the problem becomes the standard
    symmetric one K y = lam y, K_ij = Wt_i A_ij Wt_j^T, v_i = Wt_i^T y_i.  K is indefinite (for two
    views its spectrum is +-sigma), so it is solved shifted by ||K||_F to keep +-pairs apart.
    """
    m = len(dims)
    sl = _slices(dims)
    D = C.shape[0]
    lams, vts = _block_eigh(C, dims)
    # eps floor of _build_B (:170-172): lambda_min of the block-diagonal B is the min over blocks
    min_eig = min(float(((1.0 - c[i]) * lams[i][-1] + c[i]).item()) for i in range(m))
    floor = (eps - min_eig) if min_eig < eps else 0.0
    wts = []
    for i in range(m):
        Wt, _, _ = ops.whiten_rows(lams[i], vts[i], c[i], floor_add=floor, scale=1.0 / m, rank_tol=-1.0)
        wts.append(Wt)
    K = torch.zeros((D, D), dtype=C.dtype, device=C.device)
    for i in range(m):
        for j in range(i + 1, m):
            tmp = ops.gemm(wts[i], C[sl[i], sl[j]])
            ops.gemm(tmp, wts[j], transb=True, alpha=1.0 / m, out=K[sl[i], sl[j]])
            K[sl[j], sl[i]] = K[sl[i], sl[j]].T
    shift = float(ops.frobenius_norm(K).item())
    evals, evt = ops.syevj(K, shift=shift)
    k = min(latent_dimensions, D)
    return [ops.gemm(wts[i], evt[:k, sl[i]], transa=True, transb=True) for i in range(m)]


def gcca_weights(C, dims, n_samples, latent_dimensions, c, mu, eps):
    """GCCA in primal (D x D) form (cca_zoo/linear/_gcca.py:94-109).

    reg_i = (1-c_i) L_i + c_i (+ per-view eps floor, :102-104) ; Wt_i = diag(sqrt(mu_i) reg_i^-1/2) V_i^T ;
    G = (n-1) Wt C Wt^T (block-wise) ; top-k G u = sig u ;
    W_i = pinv(C_ii) [C Wt^T u]_i sig^-1/2   (pinv from the same eigendecomposition).
    """
    m = len(dims)
    sl = _slices(dims)
    D = C.shape[0]
    lams, vts = _block_eigh(C, dims)
    wts = []
    for i in range(m):
        reg_min = float(((1.0 - c[i]) * lams[i][-1] + c[i]).item())
        floor = (eps - reg_min) if reg_min < eps else 0.0
        Wt, _, _ = ops.whiten_rows(lams[i], vts[i], c[i], floor_add=floor, scale=1.0 / mu[i], rank_tol=-1.0)
        wts.append(Wt)
    G = torch.empty((D, D), dtype=C.dtype, device=C.device)
    for i in range(m):
        for j in range(i, m):
            tmp = ops.gemm(wts[i], C[sl[i], sl[j]])
            ops.gemm(tmp, wts[j], transb=True, alpha=float(n_samples - 1), out=G[sl[i], sl[j]])
            if j > i:
                G[sl[j], sl[i]] = G[sl[i], sl[j]].T
    sig, evt = ops.syevj(G)
    k = min(latent_dimensions, D, n_samples)
    # P = blkdiag(Wt_i^T) U_k   (D x k)
    P = torch.empty((D, k), dtype=C.dtype, device=C.device)
    for i in range(m):
        ops.gemm(wts[i], evt[:k, sl[i]], transa=True, transb=True, out=P[sl[i]])
    CP = ops.gemm(C, P)                                                   # (D x k)
    out = []
    for i in range(m):
        tol = _rank_tol(dims[i], C.dtype)
        # pinv(C_ii) = V diag(1/lam | lam > tol lam_max) V^T  via whiten_rows with c=0 (g = lam^-1/2) twice
        Pinv_half, _, _ = ops.whiten_rows(lams[i], vts[i], 0.0, rank_tol=tol)   # diag(lam^-1/2) V^T
        t1 = ops.gemm(Pinv_half, CP[sl[i]])                                      # (d x k)
        wi = ops.gemm(Pinv_half, t1, transa=True)                                # V lam^-1 V^T CP_i
        out.append(ops.scale(wi, cols=sig[:k], cols_pow=-0.5))
    return out
