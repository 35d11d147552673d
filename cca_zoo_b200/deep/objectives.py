"""Differentiable CCA objectives on the GPU (mirrors cca_zoo/deep/objectives.py:24-220).

``CCALoss.forward([z1, z2])`` returns ``-|| S11^-1/2 S12 S22^-1/2 ||_F^2`` with
``Sii = cov(zi) + eps I`` and eigenvalues clamped at ``eps`` exactly as the reference
(objectives.py:86-102, ``_inv_sqrtm`` :9-21).  Plug into ``DCCA(objective=...)``
(cca_zoo/deep/_dcca.py:61,73) unchanged: it is an ``nn.Module`` taking ``list[Tensor]`` and returning
a 0-dim tensor.

Forward  : K1 moments of [z1 z2] -> covariance S -> whitening of each view -> T -> loss = -||T||_F^2
           (= -sum eigvalsh(T^T T): the third eigensolve of the reference is a trace).  Whitening uses the
           Cholesky factors S_ii + eps I = L_i L_i^T (T = L_1^-1 S_12 L_2^-T, same Frobenius norm as
           S_11^-1/2 S_12 S_22^-1/2) whenever lambda_min is provably above the clamp; otherwise two Jacobi
           eigendecompositions reproduce clamp(eigh(.), min=eps) literally.
Backward : analytic, no eigh-backward:  with P = S11^-1 S12 S22^-1,
           dL/dz1 = 2/(n-1) * center(z1 (P S21 S11^-1) - z2 P^T),  dL/dz2 symmetric.
           Valid whenever the eigenvalue clamp is inactive, which ``+ eps I`` guarantees up to round-off.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops


def _require_cuda(name, *tensors):
    if not all(t.is_cuda for t in tensors):
        raise RuntimeError(f"cca_zoo_b200.{name} needs CUDA tensors (sm_100a); there is no CPU fallback.")


def _whiteners_cholesky(C, d1, eps):
    """Cholesky route: S_ii = C_ii + eps I = L_i L_i^T, returns (L_1, L_2) or None.

    Accepted only when every pivot^2 exceeds 4 eps, i.e. lambda_min(S_ii) is safely above the reference's
    eigenvalue clamp (objectives.py:20) so that the clamp is provably inactive and
    S_ii^-1/2 S_12 S_jj^-1/2 has the same Frobenius norm as L_i^-1 S_12 L_j^-T."""
    Ls, flags = [], []
    for blk in (C[:d1, :d1], C[d1:, d1:]):
        S = blk.contiguous()
        S.diagonal().add_(eps)
        flags.append(ops.potrf_(S, pivot_tol=4.0 * eps))
        Ls.append(S)
    if int(torch.stack(flags).max().item()) != 0:   # one host read-back decides the route
        return None
    return Ls


class _CCALossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z1, z2, eps, precision):
        _require_cuda("CCALoss", z1, z2)
        if z1.dtype != z2.dtype or z1.dtype not in (torch.float32, torch.float64):
            raise ValueError("representations must share a float32/float64 dtype")
        n = z1.shape[0]
        d1, d2 = z1.shape[1], z2.shape[1]
        z1d, z2d = z1.detach(), z2.detach()
        mom = ops.moments([z1d, z2d], precision=precision)
        C, _ = ops.covariance(mom, [d1, d2], n, center=True, dtype=z1.dtype)
        ctx.n = n
        if max(d1, d2) <= 64:
            # fused small-matrix stage (K6): loss and the three gradient matrices in one single-CTA launch
            loss1, G11, P, G22, minp = ops.ccaloss_small(C, d1, d2, eps)
            if float(minp.item()) > 4.0 * eps:          # clamp provably inactive (one host read-back)
                ctx.fused = True
                ctx.save_for_backward(z1d, z2d, G11, P, G22)
                return loss1.reshape(()).clone()
        ctx.fused = False
        S12 = C[:d1, d1:].contiguous()
        Ls = _whiteners_cholesky(C, d1, eps)
        if Ls is not None:
            T = S12.clone()
            ops.trsm_(Ls[0], T, side="left")                  # L1^-1 S12
            ops.trsm_(Ls[1], T, side="right", trans=True)     # ... L2^-T
            # S_ii^-1 = Linv_i^T Linv_i with Linv_i = L_i^-1 (small: d x d)
            inv = []
            for L in Ls:
                E = torch.eye(L.shape[0], dtype=L.dtype, device=L.device)
                inv.append(ops.trsm_(L, E, side="left"))
            W1t, W2t = inv
        else:
            # eigen route: reproduces clamp(eigh(S + eps I), min=eps) exactly (rank-deficient batches)
            whiten = []
            for blk in (C[:d1, :d1], C[d1:, d1:]):
                lam, Vt = ops.syevj(blk.contiguous())
                Wt, _, _ = ops.whiten_rows(lam, Vt, 0.0, floor_add=eps, rank_tol=-1.0, lam_floor=0.0)
                whiten.append(Wt)
            W1t, W2t = whiten
            T = ops.gemm(ops.gemm(W1t, S12), W2t, transb=True)
        fro = ops.frobenius_norm(T)
        loss = -(fro * fro).reshape(())
        ctx.save_for_backward(z1d, z2d, W1t, W2t, S12)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        n = ctx.n
        if ctx.fused:
            z1, z2, g11, P, g22 = ctx.saved_tensors
        else:
            z1, z2, W1t, W2t, S12 = ctx.saved_tensors
            S1inv = ops.gemm(W1t, W1t, transa=True)          # S11^-1 = W1 W1^T  (W_i^T = L_i^-1 or Lam^-1/2 V^T)
            S2inv = ops.gemm(W2t, W2t, transa=True)
            P = ops.gemm(ops.gemm(S1inv, S12), S2inv)        # d1 x d2
            g11 = ops.gemm(ops.gemm(P, S12, transb=True), S1inv)   # P S21 S11^-1
            g22 = ops.gemm(ops.gemm(S2inv, S12, transa=False, transb=True), P)  # S22^-1 S21 P
        a = 2.0 / (n - 1)
        g1 = ops.gemm(z1, g11, alpha=a)
        ops.gemm(z2, P, transb=True, alpha=-a, beta=1.0, out=g1)
        g2 = ops.gemm(z2, g22, alpha=a)
        ops.gemm(z1, P, alpha=-a, beta=1.0, out=g2)
        ops.center_columns_(g1)
        ops.center_columns_(g2)
        go = grad_out.to(g1.dtype)
        return g1 * go, g2 * go, None, None


class CCALoss(nn.Module):
    r"""Andrew et al. (2013) deep-CCA loss for two views (cca_zoo/deep/objectives.py:24-102).

    Args:
        eps: ridge added to the within-view covariances and eigenvalue floor (default 1e-5).
        precision: arithmetic of the covariance kernel for float32 inputs
            (``"exact"`` default: mini-batches are HBM/latency bound, CUDA-core FMA is free).
    """

    def __init__(self, eps: float = 1e-5, precision: str = "exact") -> None:
        super().__init__()
        self.eps = eps
        self.precision = precision

    def forward(self, representations: list[torch.Tensor]) -> torch.Tensor:
        if len(representations) != 2:
            raise ValueError(
                "CCALoss expects exactly 2 representations, "
                f"got {len(representations)}."
            )
        z1, z2 = representations
        return _CCALossFn.apply(z1, z2, float(self.eps), self.precision)


class MCCALoss(nn.Module):
    r"""Sum of pairwise CCA losses over all view pairs (cca_zoo/deep/objectives.py:105-153)."""

    def __init__(self, eps: float = 1e-5, precision: str = "exact") -> None:
        super().__init__()
        self.eps = eps
        self._cca_loss = CCALoss(eps=eps, precision=precision)

    def forward(self, representations: list[torch.Tensor]) -> torch.Tensor:
        n_views = len(representations)
        total = torch.zeros((), device=representations[0].device, dtype=representations[0].dtype)
        for i in range(n_views):
            for j in range(i + 1, n_views):
                total = total + self._cca_loss([representations[i], representations[j]])
        return total


class _GCCALossFn(torch.autograd.Function):
    """MAX-VAR GCCA objective in its primal form.

    The reference (objectives.py:196-220) builds the n x n matrix M = sum_i H_i H_i^T, H_i = Zc_i S_i^-1/2, and
    sums its top-k eigenvalues: O(n^2) memory, O(n^3) work per step.  With H = [H_1 .. H_m] (n x D) the non-zero
    spectrum of M = H H^T equals that of K = H^T H = (n-1) Wt C Wt^T (D x D), Wt = blkdiag(Wt_i) with
    Wt_i = diag(clamp(lam_i + eps, min=eps)^-1/2) V_i^T -- all of it a function of the block covariance C that
    one K1 pass provides.  Backward is analytic (Hellmann-Feynman on the eigenvalue sum, no eigh-backward):
        Q = Wt^T U_k Lam_k^-1/2,  A = (n-1) C Q,  B = blkdiag(S_i^-1) A,
        dL/dz_i = center( -2 (Z Q) B_i^T + 2/(n-1) z_i B_i B_i^T ).
    """

    @staticmethod
    def forward(ctx, eps, precision, *zs):
        _require_cuda("GCCALoss", *zs)
        dt = zs[0].dtype
        if dt not in (torch.float32, torch.float64) or any(z.dtype != dt for z in zs):
            raise ValueError("representations must share a float32/float64 dtype")
        n = zs[0].shape[0]
        dims = [int(z.shape[1]) for z in zs]
        D, k = sum(dims), dims[0]
        zd = [z.detach() for z in zs]
        mom = ops.moments(zd, precision=precision)
        C, _ = ops.covariance(mom, dims, n, center=True, dtype=dt)
        Wt = torch.zeros((D, D), dtype=dt, device=C.device)
        off = 0
        for d in dims:
            lam, Vt = ops.syevj(C[off:off + d, off:off + d].contiguous())
            Wi, _, _ = ops.whiten_rows(lam, Vt, 0.0, floor_add=eps, rank_tol=-1.0, lam_floor=0.0)
            Wt[off:off + d, off:off + d] = Wi
            off += d
        K = ops.gemm(ops.gemm(Wt, C), Wt, transb=True, alpha=float(n - 1))
        K = 0.5 * (K + K.T)
        evals, Ut = ops.syevj(K)                              # descending; rows of Ut are eigenvectors
        lam_k = evals[:k].contiguous()
        ctx.n, ctx.dims = n, dims
        ctx.save_for_backward(C, Wt, lam_k, Ut[:k].contiguous(), *zd)
        return -lam_k.sum()

    @staticmethod
    def backward(ctx, grad_out):
        C, Wt, lam_k, Ut_k, *zs = ctx.saved_tensors
        n, dims = ctx.n, ctx.dims
        k = lam_k.shape[0]
        safe = lam_k.clamp_min(torch.finfo(lam_k.dtype).tiny * 1e8)
        Qt = ops.scale(Ut_k, rows=safe, rows_pow=-0.5)                    # k x D : Lam^-1/2 U_k^T
        Q = ops.gemm(Wt, Qt, transa=True, transb=True)                    # D x k
        A = ops.gemm(C, Q, alpha=float(n - 1))                            # D x k
        B = ops.gemm(Wt, ops.gemm(Wt, A), transa=True)                    # blkdiag(S_i^-1) A
        Y = torch.zeros((n, k), dtype=C.dtype, device=C.device)
        off = 0
        for z, d in zip(zs, dims):
            ops.gemm(z, Q[off:off + d], beta=1.0, out=Y)
            off += d
        go = grad_out.to(C.dtype)
        grads, off = [], 0
        for z, d in zip(zs, dims):
            Bi = B[off:off + d]
            g = ops.gemm(Y, Bi, transb=True, alpha=-2.0)
            ops.gemm(z, ops.gemm(Bi, Bi, transb=True), alpha=2.0 / (n - 1), beta=1.0, out=g)
            ops.center_columns_(g)
            grads.append(g * go)
            off += d
        return (None, None, *grads)


class GCCALoss(nn.Module):
    r"""Generalised (MAX-VAR) CCA loss for two or more views (cca_zoo/deep/objectives.py:156-220):
    :math:`-\sum_{d \le k} \lambda_d(\sum_i H_i H_i^\top)` with ``k`` the width of the first representation.
    Same constructor and ``forward(list[Tensor]) -> 0-dim Tensor`` as the reference, so it plugs into
    ``DCCA(objective=...)`` and is what ``DGCCA`` uses (cca_zoo/deep/_dgcca.py:70).  At most 8 views."""

    def __init__(self, eps: float = 1e-5, precision: str = "exact") -> None:
        super().__init__()
        self.eps = eps
        self.precision = precision

    def forward(self, representations: list[torch.Tensor]) -> torch.Tensor:
        return _GCCALossFn.apply(float(self.eps), self.precision, *representations)
