"""Differentiable CCA objectives on the GPU (mirrors cca_zoo/deep/objectives.py:24-220).

``CCALoss.forward([z1, z2])`` returns ``-|| S11^-1/2 S12 S22^-1/2 ||_F^2`` with
``Sii = cov(zi) + eps I`` and eigenvalues clamped at ``eps`` exactly as the reference
(objectives.py:86-102, ``_inv_sqrtm`` :9-21).  Plug into ``DCCA(objective=...)``
(cca_zoo/deep/_dcca.py:61,73) unchanged: it is an ``nn.Module`` taking ``list[Tensor]`` and returning
a 0-dim tensor.

Forward  : ONE library call (``ccab_ccaloss_fwd``, csrc/fit.cu): moment pass over [z1 z2] -> S -> batched Cholesky +
           explicit inverse of S_11, S_22 -> P = S11^-1 S12 S22^-1 and the two small gradient matrices ->
           loss = -<P, S12> (the reference's third eigensolve, eigvalsh(T^T T).sum(), is a trace).  Nothing is read
           back on this path: the Cholesky status lands in a device flag that is checked LAZILY (at the next call, or
           by ``check()``).  Batches that are rank deficient by shape (n - 1 < width: the eigenvalue clamp of the
           reference is then active for certain) take the eigen route, which reproduces
           ``clamp(eigh(S + eps I), min=eps)`` literally with two Jacobi eigendecompositions.
Backward : analytic, no eigh-backward, ONE library call (``ccab_ccaloss_bwd``): with
           P = S11^-1 S12 S22^-1,  dL/dz1 = 2/(n-1) * center(z1 (P S21 S11^-1) - z2 P^T),  dL/dz2 symmetric.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops


def _require_cuda(name, *tensors):
    if not all(t.is_cuda for t in tensors):
        raise RuntimeError(f"cca_zoo_b200.{name} needs CUDA tensors (sm_100a); there is no CPU fallback.")


def _row_major(z):
    z = z.detach()
    return z if (z.stride(1) == 1 and z.stride(0) >= z.shape[1]) else z.contiguous()


def _resolve_precision(precision, zs):
    """"auto": exact FMA moments for narrow batches (HBM / latency bound), the tcgen05 3xTF32 kernel once the block
    covariance is wide enough to be a real contraction -- if TMA can address the representations."""
    if zs[0].dtype == torch.float64:
        return "exact"
    if precision == "auto":
        precision = "tf32x3b" if sum(z.shape[1] for z in zs) > 256 else "exact"
    if precision != "exact" and not all(z.data_ptr() % 16 == 0 and z.stride(0) % 4 == 0 for z in zs):
        return "exact"
    return precision


class _LazyStatus:
    """Device-side status flags of past evaluations, copied to pinned memory asynchronously and inspected without
    ever blocking the stream: ``poll`` looks at the copies that have already landed, ``check`` waits for all."""

    def __init__(self, what, nan_last=True):
        self.what = what
        self.nan_last = nan_last       # the last flag reports NaN / inf in the input
        self.pending = []
        self.free = []                 # recycled (pinned buffer, event) pairs: cudaHostAlloc costs ~0.1 ms

    def push(self, flags):
        if not flags.is_cuda:          # host-logic tests (tests/fake_ops.py): nothing is asynchronous there
            self._inspect(flags)
            return
        n = flags.numel()
        slot = next((i for i, (h, _) in enumerate(self.free) if h.numel() >= n), None)
        if slot is None:
            buf, ev = torch.empty(max(n, 16), dtype=flags.dtype, pin_memory=True), torch.cuda.Event()
        else:
            buf, ev = self.free.pop(slot)
        host = buf[:n]
        host.copy_(flags, non_blocking=True)
        ev.record(torch.cuda.current_stream(flags.device))
        self.pending.append((host, ev, buf))

    def _inspect(self, host):
        vals = host.tolist()
        if any(v != 0 for v in vals):
            self.pending.clear()
            if self.nan_last and vals[-1]:
                raise ValueError(f"{self.what}: a representation contained NaN or infinity.")
            raise RuntimeError(
                f"{self.what}: a within-view covariance S_ii + eps I of an earlier batch was not numerically positive "
                f"definite (Cholesky status {vals}); that loss value and its gradients are unreliable.  Use "
                f"verify='sync' to take the eigen route for such batches automatically, or a larger eps.")

    def poll(self):
        while self.pending and self.pending[0][1].query():
            host, ev, buf = self.pending.pop(0)
            self.free.append((buf, ev))
            self._inspect(host)

    def check(self):
        while self.pending:
            host, ev, buf = self.pending.pop(0)
            ev.synchronize()
            self.free.append((buf, ev))
            self._inspect(host)


def _eigen_route(z1d, z2d, eps, precision):
    """clamp(eigh(S + eps I), min=eps) literally (objectives.py:19-21): rank-deficient batches, verify='sync' fallback."""
    n, d1 = z1d.shape[0], z1d.shape[1]
    mom = ops.moments([z1d, z2d], precision=precision)
    C, _ = ops.covariance(mom, [d1, z2d.shape[1]], n, center=True, dtype=z1d.dtype)
    S12 = C[:d1, d1:].contiguous()
    whiten = []
    for blk in (C[:d1, :d1], C[d1:, d1:]):
        lam, Vt = ops.syevj(blk.contiguous())
        Wt, _, _ = ops.whiten_rows(lam, Vt, 0.0, floor_add=eps, rank_tol=-1.0, lam_floor=0.0)
        whiten.append(Wt)
    W1t, W2t = whiten
    T = ops.gemm(ops.gemm(W1t, S12), W2t, transb=True)
    fro = ops.frobenius_norm(T)
    loss = -(fro * fro)
    S1inv = ops.gemm(W1t, W1t, transa=True)              # S11^-1 = W1 W1^T (W_i^T = Lam^-1/2 V^T)
    S2inv = ops.gemm(W2t, W2t, transa=True)
    P = ops.gemm(ops.gemm(S1inv, S12), S2inv)
    g11 = ops.gemm(ops.gemm(P, S12, transb=True), S1inv)
    g22 = ops.gemm(ops.gemm(S2inv, S12, transb=True), P)
    means = torch.cat([z1d.mean(dim=0), z2d.mean(dim=0)])     # the fused narrow backward centres algebraically
    saved = torch.cat([g11.reshape(-1), P.reshape(-1), g22.reshape(-1), means])
    return loss, saved


class _CCALossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z1, z2, eps, precision, status, sync):
        _require_cuda("CCALoss", z1, z2)
        if z1.dtype != z2.dtype or z1.dtype not in (torch.float32, torch.float64):
            raise ValueError("representations must share a float32/float64 dtype")
        n = z1.shape[0]
        z1d, z2d = _row_major(z1), _row_major(z2)
        prec = _resolve_precision(precision, [z1d, z2d])
        if n - 1 < max(z1d.shape[1], z2d.shape[1]):          # rank deficient by shape: the clamp is active for certain
            loss, saved = _eigen_route(z1d, z2d, eps, prec)
        else:
            loss, saved, flags = ops.ccaloss_fwd(z1d, z2d, eps, prec)
            if sync:
                f = flags.tolist()                            # verify='sync': one read-back per step
                if f[2]:
                    raise ValueError("CCALoss: a representation contained NaN or infinity.")
                if f[0] or f[1]:
                    loss, saved = _eigen_route(z1d, z2d, eps, prec)
            else:
                status.push(flags)
        ctx.save_for_backward(z1d, z2d, saved)
        return loss.reshape(()).clone()

    @staticmethod
    def backward(ctx, grad_out):
        z1, z2, saved = ctx.saved_tensors
        go = grad_out.to(z1.dtype).reshape(1).contiguous()
        g1, g2 = ops.ccaloss_bwd(z1, z2, saved, go)
        return g1, g2, None, None, None, None


class CCALoss(nn.Module):
    r"""Andrew et al. (2013) deep-CCA loss for two views (cca_zoo/deep/objectives.py:24-102).

    Args:
        eps: ridge added to the within-view covariances and eigenvalue floor (default 1e-5).
        precision: arithmetic of the covariance kernel for float32 inputs: ``"auto"`` (default: exact CUDA-core FMA
            for narrow representations, tcgen05 3xTF32 beyond a total width of 256), ``"exact"``, ``"tf32x3"``, ``"tf32"``.
        verify: ``"lazy"`` (default) never reads anything back in ``forward``: the Cholesky status of every
            evaluation is copied to the host asynchronously and inspected at the next call / by ``check()``, which
            raise if an earlier batch had a numerically indefinite covariance.  ``"sync"`` reads the status back in
            every call and takes the eigen route (the reference's ``clamp(eigh(.))`` literally) for such batches.
    """

    def __init__(self, eps: float = 1e-5, precision: str = "auto", verify: str = "lazy") -> None:
        super().__init__()
        if verify not in ("lazy", "sync"):
            raise ValueError("verify must be 'lazy' or 'sync'")
        self.eps = eps
        self.precision = precision
        self.verify = verify
        self._status = _LazyStatus("CCALoss")

    def check(self) -> None:
        """Wait for the status of every evaluation issued so far and raise if one of them was unreliable."""
        self._status.check()

    def forward(self, representations: list[torch.Tensor]) -> torch.Tensor:
        if len(representations) != 2:
            raise ValueError(
                "CCALoss expects exactly 2 representations, "
                f"got {len(representations)}."
            )
        self._status.poll()
        z1, z2 = representations
        return _CCALossFn.apply(z1, z2, float(self.eps), self.precision, self._status, self.verify == "sync")


class _MCCALossFn(torch.autograd.Function):
    """Sum of the pairwise CCA losses from ONE moment pass over all views, every S_ii factored ONCE
    (the reference calls CCALoss per pair, objectives.py:148-153: m - 1 eigendecompositions of every S_ii and m - 1
    passes over every z_i).  With A_i = S_ii^-1 (batched Cholesky + inverse), P_ij = A_i S_ij A_j:
        loss = - sum_{i<j} <P_ij, S_ij>,
        dL/dz_i = 2/(n-1) center( z_i G_i - sum_{j != i} z_j P_ji ),   G_i = sum_{j != i} P_ij S_ji A_i,  P_ji = P_ij^T.
    """

    @staticmethod
    def forward(ctx, eps, precision, status, *zs):
        _require_cuda("MCCALoss", *zs)
        dt = zs[0].dtype
        if dt not in (torch.float32, torch.float64) or any(z.dtype != dt for z in zs):
            raise ValueError("representations must share a float32/float64 dtype")
        n, m = zs[0].shape[0], len(zs)
        zd = [_row_major(z) for z in zs]
        dims = [int(z.shape[1]) for z in zd]
        off = [0]
        for d in dims:
            off.append(off[-1] + d)
        mom = ops.moments(zd, precision=_resolve_precision(precision, zd))
        C, _ = ops.covariance(mom, dims, n, center=True, dtype=dt)
        A, flags = [], []
        if len(set(dims)) == 1:                                  # one batched factorisation for all views
            d = dims[0]
            R = torch.stack([C[off[i]:off[i + 1], off[i]:off[i + 1]] for i in range(m)])
            R.diagonal(dim1=1, dim2=2).add_(eps)
            Linv, info = ops.potrf_inv_(R, pivot_tol=0.25 * eps)
            flags.append(info)
            Ab = ops.gemm_batched(Linv, Linv, transa=True)
            A = [Ab[i] for i in range(m)]
        else:
            for i in range(m):
                R = C[off[i]:off[i + 1], off[i]:off[i + 1]].contiguous()
                R.diagonal().add_(eps)
                Linv, info = ops.potrf_inv_(R, pivot_tol=0.25 * eps)
                flags.append(info)
                A.append(ops.gemm(Linv, Linv, transa=True))
        G = [torch.zeros((d, d), dtype=dt, device=C.device) for d in dims]
        P = {}
        terms = []
        for i in range(m):
            for j in range(i + 1, m):
                Sij = C[off[i]:off[i + 1], off[j]:off[j + 1]]
                Q = ops.gemm(A[i], Sij)                          # A_i S_ij
                Q2 = ops.gemm(Sij, A[j])                         # S_ij A_j
                Pij = ops.gemm(Q, A[j])
                ops.gemm(Pij, Q, transb=True, beta=1.0, out=G[i])            # += P_ij S_ji A_i
                ops.gemm(Q2, Pij, transa=True, beta=1.0, out=G[j])           # += A_j S_ji P_ij
                P[(i, j)] = Pij
                terms.append((Pij * Sij).sum())
        status.push(torch.cat(flags))
        ctx.n, ctx.m = n, m
        ctx.pairs = sorted(P)
        ctx.save_for_backward(*zd, *G, *[P[k] for k in ctx.pairs])
        return -torch.stack(terms).sum()

    @staticmethod
    def backward(ctx, grad_out):
        m, n = ctx.m, ctx.n
        t = ctx.saved_tensors
        zs, G = t[:m], t[m:2 * m]
        P = dict(zip(ctx.pairs, t[2 * m:]))
        a = 2.0 / (n - 1)
        go = grad_out.to(zs[0].dtype)
        grads = []
        for i in range(m):
            g = ops.gemm(zs[i], G[i], alpha=a)
            for j in range(m):
                if j == i:
                    continue
                if i < j:
                    ops.gemm(zs[j], P[(i, j)], transb=True, alpha=-a, beta=1.0, out=g)   # - z_j P_ij^T
                else:
                    ops.gemm(zs[j], P[(j, i)], alpha=-a, beta=1.0, out=g)                # - z_j P_ji
            ops.center_columns_(g)
            grads.append(g * go)
        return (None, None, None, *grads)


class MCCALoss(nn.Module):
    r"""Sum of pairwise CCA losses over all view pairs (cca_zoo/deep/objectives.py:105-153), computed from one
    moment pass with every within-view covariance factored once.  Same ``verify`` semantics as ``CCALoss``
    (``"lazy"``: status checked at the next call; ``"sync"``: per-pair ``CCALoss`` evaluations with the eigen-route
    fallback, i.e. the reference's loop)."""

    def __init__(self, eps: float = 1e-5, precision: str = "auto", verify: str = "lazy") -> None:
        super().__init__()
        if verify not in ("lazy", "sync"):
            raise ValueError("verify must be 'lazy' or 'sync'")
        self.eps = eps
        self.precision = precision
        self.verify = verify
        self._status = _LazyStatus("MCCALoss", nan_last=False)
        self._cca_loss = CCALoss(eps=eps, precision=precision, verify="sync")

    def check(self) -> None:
        self._status.check()

    def forward(self, representations: list[torch.Tensor]) -> torch.Tensor:
        n_views = len(representations)
        n = representations[0].shape[0]
        lazy_ok = (self.verify == "lazy" and 2 <= n_views <= 8
                   and n - 1 >= max(int(z.shape[1]) for z in representations))
        if lazy_ok:
            self._status.poll()
            return _MCCALossFn.apply(float(self.eps), self.precision, self._status, *representations)
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
        prec = _resolve_precision(self.precision, [_row_major(z) for z in representations])
        return _GCCALossFn.apply(float(self.eps), prec, *representations)
