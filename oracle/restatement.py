"""CPU restatement (numpy/scipy, float64 unless told otherwise) of the reference hot path.

TEST INFRASTRUCTURE ONLY -- this module is the *checker* for the CUDA path.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product (``cca_zoo_b200``) never does.

Parity status: PINNED.  Every function below is checked against the unmodified
reference (imported through ``oracle/refshim.py``) by ``tests/test_oracle_vs_reference.py``
in the authoring container, and against the committed fixtures in ``tests/golden``
(made by ``oracle/make_golden.py`` from the reference itself) everywhere else.

Two families of functions:

* ``ref_*``  -- line-by-line restatements of the reference algorithms (tall SVD, n x n
  GCCA matrix, eigh-based loss).  These *are* the reference's arithmetic.
* ``cov_*``  -- the covariance-space forms the CUDA kernels implement:
  everything is a function of the block moment matrix ``M = [X1..Xm]^T [X1..Xm]``,
  the column sums ``s`` and ``n``.  They are validated against ``ref_*``.

Reference citations are relative to /root/reference.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg

# --------------------------------------------------------------------------------------
# shared pieces
# --------------------------------------------------------------------------------------


def perview(value, default, m):
    """cca_zoo/_utils/_validation.py:45-75 (perview_parameter)."""
    if value is None:
        return [default] * m
    if isinstance(value, (list, tuple)):
        if len(value) != m:
            raise ValueError("per-view parameter has wrong length")
        return list(value)
    return [value] * m


def setup_fit(views, center=True):
    """cca_zoo/_base.py:78-102 -- means and centred copies (dtype preserved)."""
    views = [np.asarray(v) for v in views]
    if center:
        means = [v.mean(axis=0) for v in views]
        views = [v - mu for v, mu in zip(views, means)]
    else:
        means = [np.zeros(v.shape[1]) for v in views]
    return views, means


def transform(views, means, weights):
    """cca_zoo/_base.py:108-123."""
    return [(np.asarray(v) - mu) @ w for v, mu, w in zip(views, means, weights)]


def pairwise_correlations(variates):
    """cca_zoo/_base.py:153-174 on already-transformed variates."""
    T = np.stack(variates, axis=0)
    T = T - T.mean(axis=1, keepdims=True)
    norms = np.sqrt((T**2).sum(axis=1, keepdims=True))
    Tn = T / np.where(norms > 1e-12, norms, 1.0)
    return np.einsum("isd,jsd->ijd", Tn, Tn)


def average_pairwise_correlations(variates):
    """cca_zoo/_base.py:176-194."""
    corrs = pairwise_correlations(variates)
    m = corrs.shape[0]
    off = corrs.sum(axis=(0, 1)) - sum(corrs[i, i, :] for i in range(m))
    return off / (m * (m - 1))


def score(views, means, weights):
    """cca_zoo/_base.py:140-151."""
    return average_pairwise_correlations(transform(views, means, weights))


# --------------------------------------------------------------------------------------
# ref_* : the reference's own arithmetic
# --------------------------------------------------------------------------------------


def ref_svd_whiten(X, regularization=0.0):
    """cca_zoo/_utils/_linalg.py:9-41."""
    n = X.shape[0]
    U, s, Vt = np.linalg.svd(X, full_matrices=False)
    pos = s > 0
    s, U, Vt = s[pos], U[:, pos], Vt[pos, :]
    lam = s**2 / (n - 1)
    inv_sqrt = 1.0 / np.sqrt((1.0 - regularization) * lam + regularization)
    return U * (s * inv_sqrt), Vt.T * inv_sqrt


def ref_gevp(A, B, k):
    """cca_zoo/_utils/_linalg.py:44-73."""
    p = A.shape[0]
    kc = min(k, p)
    if B is None:
        w, v = scipy.linalg.eigh(A, subset_by_index=[p - kc, p - 1])
    else:
        w, v = scipy.linalg.eigh(A, B, subset_by_index=[p - kc, p - 1])
    idx = np.argsort(w)[::-1]
    return w[idx].real, v[:, idx].real


def ref_rcca_fit(views, latent_dimensions=1, c=0.0, center=True):
    """cca_zoo/linear/_rcca.py:69-101.  Returns (weights, means)."""
    vs, means = setup_fit(views, center)
    if len(vs) != 2:
        raise ValueError("rCCA requires exactly 2 views")
    c_ = perview(c, 0.0, 2)
    X1, X2 = vs
    X1w, W1 = ref_svd_whiten(X1, c_[0])
    X2w, W2 = ref_svd_whiten(X2, c_[1])
    k = min(latent_dimensions, X1w.shape[1], X2w.shape[1])
    cross = X1w.T @ X2w / (X1.shape[0] - 1)
    U, _, Vt = np.linalg.svd(cross, full_matrices=False)
    return [W1 @ U[:, :k], W2 @ Vt[:k, :].T], means


def ref_mcca_fit(views, latent_dimensions=1, c=0.0, eps=1e-6, center=True):
    """cca_zoo/linear/_mcca.py:99-197 with pca=False (``_build_A`` :141-153,
    ``_build_B`` :155-173).  ``pca=True`` gives the same weights for full-column-rank
    views (checked against the reference in tests/test_oracle_vs_reference.py)."""
    vs, means = setup_fit(views, center)
    m = len(vs)
    c_ = perview(c, 0.0, m)
    A = np.cov(np.hstack(vs), rowvar=False)
    A = A - scipy.linalg.block_diag(*[np.atleast_2d(np.cov(v, rowvar=False)) for v in vs])
    A = A / m
    blocks = [
        (1.0 - c_[i]) * np.atleast_2d(np.cov(v, rowvar=False)) + c_[i] * np.eye(v.shape[1])
        for i, v in enumerate(vs)
    ]
    B = scipy.linalg.block_diag(*blocks)
    min_eig = np.linalg.eigvalsh(B).min()
    if min_eig < eps:
        B = B + (eps - min_eig) * np.eye(B.shape[0])
    B = B / m
    _, vecs = ref_gevp(A, B, latent_dimensions)
    splits = np.cumsum([v.shape[1] for v in vs])
    return np.split(vecs, splits[:-1], axis=0), means


def ref_gcca_fit(views, latent_dimensions=1, c=0.0, view_weights=None, eps=1e-6, center=True):
    """cca_zoo/linear/_gcca.py:80-110 (forms the n x n matrix: small n only)."""
    vs, means = setup_fit(views, center)
    m = len(vs)
    n = vs[0].shape[0]
    c_ = perview(c, 0.0, m)
    mu = perview(view_weights, 1.0, m)
    Q = np.zeros((n, n))
    for v, ci, mi in zip(vs, c_, mu):
        cov_i = (1.0 - ci) * np.atleast_2d(np.cov(v, rowvar=False)) + ci * np.eye(v.shape[1])
        min_eig = np.linalg.eigvalsh(cov_i).min()
        if min_eig < eps:
            cov_i = cov_i + (eps - min_eig) * np.eye(cov_i.shape[0])
        Q += mi * (v @ np.linalg.inv(cov_i) @ v.T)
    _, vecs = ref_gevp(Q, None, latent_dimensions)
    T = vecs[:, :latent_dimensions]
    return [np.linalg.pinv(v) @ T for v in vs], means


def _mcca_core(vs, latent_dimensions, c_, eps):
    """A, B of cca_zoo/linear/_mcca.py:141-173 (pca=False) from already processed views + gevp."""
    m = len(vs)
    A = np.cov(np.hstack(vs), rowvar=False)
    A = A - scipy.linalg.block_diag(*[np.atleast_2d(np.cov(v, rowvar=False)) for v in vs])
    A = A / m
    blocks = [(1.0 - c_[i]) * np.atleast_2d(np.cov(v, rowvar=False)) + c_[i] * np.eye(v.shape[1])
              for i, v in enumerate(vs)]
    B = scipy.linalg.block_diag(*blocks)
    min_eig = np.linalg.eigvalsh(B).min()
    if min_eig < eps:
        B = B + (eps - min_eig) * np.eye(B.shape[0])
    B = B / m
    _, vecs = ref_gevp(A, B, latent_dimensions)
    splits = np.cumsum([v.shape[1] for v in vs])
    return np.split(vecs, splits[:-1], axis=0)


def ref_partialcca_fit(views, partials, latent_dimensions=1, c=0.0, eps=1e-6, center=True):
    """cca_zoo/linear/_partialcca.py:67-103.  Returns (weights, means, confound_betas)."""
    vs, means = setup_fit(views, center)
    P = np.asarray(partials, dtype=float)
    betas = [np.linalg.pinv(P) @ v for v in vs]
    dec = [v - P @ b for v, b in zip(vs, betas)]
    c_ = perview(c, 0.0, len(vs))
    return _mcca_core(dec, latent_dimensions, c_, eps), means, betas


def grcca_maps(dims, groups, c_, mu_):
    """The linear maps T_i (d_i x d_i') behind GRCCA's feature augmentation and weight collapse
    (cca_zoo/linear/_grcca.py:126-160): processed_i = X_i T_i, weights_i = T_i block_i.
    T_i = [ (I - P_g)/c | E diag(1/sqrt(mu_eff * counts)) ] with E the feature->group indicator and
    P_g = E diag(1/counts) E^T; T_i = I when c_i <= 0."""
    maps = []
    for d, g, ci, mi in zip(dims, groups, c_, mu_):
        if ci <= 0:
            maps.append(np.eye(d))
            continue
        ids, inv, counts = np.unique(np.asarray(g), return_inverse=True, return_counts=True)
        E = np.zeros((d, len(ids)))
        E[np.arange(d), inv] = 1.0
        mu_eff = 1.0 if mi == 0 else mi
        T1 = (np.eye(d) - E @ np.diag(1.0 / counts) @ E.T) / ci
        T2 = E @ np.diag(1.0 / np.sqrt(mu_eff * counts))
        maps.append(np.hstack([T1, T2]))
    return maps


def ref_grcca_fit(views, feature_groups, latent_dimensions=1, c=0.0, mu=0.0, eps=1e-6, center=True):
    """cca_zoo/linear/_grcca.py:76-160 restated through the linear maps of ``grcca_maps`` (the reference
    builds the same augmented views with per-group means; equality is asserted against the live reference in
    tests/test_oracle_vs_reference.py)."""
    vs, means = setup_fit(views, center)
    m = len(vs)
    c_ = perview(c, 0.0, m)
    mu_ = perview(mu, 0.0, m)
    if feature_groups is None:
        feature_groups = [np.ones(v.shape[1], dtype=int) for v in vs]
    maps = grcca_maps([v.shape[1] for v in vs], feature_groups, c_, mu_)
    processed = [v @ T for v, T in zip(vs, maps)]
    blocks = _mcca_core(processed, latent_dimensions, c_, eps)
    return [T @ b for T, b in zip(maps, blocks)], means


def cov_partialcca(M, s, n, dims, q, latent_dimensions=1, c=0.0, eps=1e-6, center=True):
    """Covariance form of PartialCCA from the moments of [X_1 .. X_m, P] (P: the q confound columns, LAST):
    beta = (P^T P)^+ P^T Xc ; cov(deconfounded) = (Xc^T Xc - G^T beta - t t^T / n)/(n-1) with
    G = P^T Xc, t = 1^T (Xc - P beta); then the MCCA solve (pca=False)."""
    D = int(sum(dims))
    Mxx, Mxp, Mpp = M[:D, :D], M[:D, D:], M[D:, D:]
    sx, sp = s[:D], s[D:]
    if center:
        mu = sx / n
        XcXc = Mxx - n * np.outer(mu, mu)
        G = Mxp.T - np.outer(sp, mu)
        colsum_xc = np.zeros(D)
    else:
        XcXc, G, colsum_xc = Mxx, Mxp.T, sx
    beta = np.linalg.pinv(Mpp) @ G
    t = colsum_xc - sp @ beta
    Cd = (XcXc - G.T @ beta - np.outer(t, t) / n) / (n - 1)
    w, _ = cov_mcca_fit(Cd, dims, latent_dimensions, c, eps)
    sl = block_slices(dims)
    return w, [beta[:, sl_i] for sl_i in sl]


def cov_grcca(C, dims, feature_groups, latent_dimensions=1, c=0.0, mu=0.0, eps=1e-6):
    """Covariance form of GRCCA: C' = T^T C T with T = blkdiag(T_i) (no extra pass over the data), MCCA solve
    on the augmented widths, weights_i = T_i block_i."""
    m = len(dims)
    c_ = perview(c, 0.0, m)
    mu_ = perview(mu, 0.0, m)
    if feature_groups is None:
        feature_groups = [np.ones(d, dtype=int) for d in dims]
    maps = grcca_maps(dims, feature_groups, c_, mu_)
    T = scipy.linalg.block_diag(*maps)
    Cp = T.T @ C @ T
    blocks, _ = cov_mcca_fit(Cp, [Tm.shape[1] for Tm in maps], latent_dimensions, c_, eps)
    return [Tm @ b for Tm, b in zip(maps, blocks)]


def ref_inv_sqrtm(A, eps=1e-5):
    """cca_zoo/deep/objectives.py:9-21."""
    L, V = np.linalg.eigh(A)
    L = np.maximum(L, eps)
    return (V / np.sqrt(L)) @ V.T


def ref_ccaloss(z1, z2, eps=1e-5):
    """cca_zoo/deep/objectives.py:79-102 (forward only, numpy)."""
    n = z1.shape[0]
    d1, d2 = z1.shape[1], z2.shape[1]
    z1 = z1 - z1.mean(axis=0)
    z2 = z2 - z2.mean(axis=0)
    s11 = z1.T @ z1 / (n - 1) + eps * np.eye(d1)
    s22 = z2.T @ z2 / (n - 1) + eps * np.eye(d2)
    s12 = z1.T @ z2 / (n - 1)
    t = ref_inv_sqrtm(s11, eps) @ s12 @ ref_inv_sqrtm(s22, eps)
    ev = np.linalg.eigvalsh(t.T @ t)
    return -np.maximum(ev, 0.0).sum()


def ref_ccaloss_torch_fwdbwd(z1, z2, eps=1e-5):
    """cca_zoo/deep/objectives.py:9-21,79-102 restated with torch on the CPU, forward AND autograd backward: the
    arithmetic (eigh-based inverse square roots, eigvalsh of T^T T, torch autograd through them) the reference runs
    in a training step.  ``z1``, ``z2``: torch CPU tensors.  Returns (loss, grad1, grad2).  Used as the CPU arm of
    bench.py's config-3 workloads and checked against tests/golden (reference outputs) by tests/test_oracle_golden."""
    import torch

    def inv_sqrtm(A):
        L, V = torch.linalg.eigh(A)
        L = torch.clamp(L, min=eps)
        return V @ torch.diag(1.0 / torch.sqrt(L)) @ V.T

    a = z1.detach().clone().requires_grad_(True)
    b = z2.detach().clone().requires_grad_(True)
    n, d1, d2 = a.shape[0], a.shape[1], b.shape[1]
    x1 = a - a.mean(dim=0)
    x2 = b - b.mean(dim=0)
    s11 = (x1.T @ x1) / (n - 1) + eps * torch.eye(d1, dtype=a.dtype)
    s22 = (x2.T @ x2) / (n - 1) + eps * torch.eye(d2, dtype=a.dtype)
    s12 = (x1.T @ x2) / (n - 1)
    t = inv_sqrtm(s11) @ s12 @ inv_sqrtm(s22)
    loss = -torch.clamp(torch.linalg.eigvalsh(t.T @ t), min=0.0).sum()
    loss.backward()
    return loss.detach(), a.grad, b.grad


def ref_gccaloss(zs, eps=1e-5):
    """cca_zoo/deep/objectives.py:196-220 (forward only; forms the n x n matrix like the reference)."""
    n = zs[0].shape[0]
    M = np.zeros((n, n))
    for z in zs:
        zc = z - z.mean(axis=0)
        cov = zc.T @ zc / (n - 1) + eps * np.eye(z.shape[1])
        h = zc @ ref_inv_sqrtm(cov, eps)
        M += h @ h.T
    ev = np.linalg.eigvalsh(M)
    return -ev[-zs[0].shape[1]:].sum()


def cov_gccaloss(zs, eps=1e-5):
    """Primal form the CUDA path implements: with H = [H_1 .. H_m] the n x n matrix is H H^T and its non-zero
    eigenvalues are those of K = H^T H = (n-1) Wt C Wt^T (D x D, D = sum of widths; Wt_i any matrix with
    Wt_i^T Wt_i = (C_ii + eps I)^-1).  Returns (loss, grads) with the analytic gradient
        Q = Wt^T U_k Lam_k^-1/2,  A_i = (n-1) C[i,:] Q,  B_i = S_i^-1 A_i,
        dL/dz_i = center( -2 (Zc Q) B_i^T + 2/(n-1) Zc_i B_i B_i^T )
    (Hellmann-Feynman on the top-k eigenvalue sum; valid when lambda_k > lambda_k+1 and the clamp is inactive)."""
    n = zs[0].shape[0]
    dims = [z.shape[1] for z in zs]
    k = dims[0]
    sl = block_slices(dims)
    Zc = np.hstack([z - z.mean(axis=0) for z in zs])
    C = Zc.T @ Zc / (n - 1)
    D = C.shape[0]
    Wt = np.zeros((D, D))
    Sinv = []
    for s_ in sl:
        lam, V = np.linalg.eigh(C[s_, s_] + eps * np.eye(s_.stop - s_.start))
        lam = np.maximum(lam, eps)
        Wt[s_, s_] = (V / np.sqrt(lam)).T
        Sinv.append((V / lam) @ V.T)
    K = (n - 1) * Wt @ C @ Wt.T
    ev, U = np.linalg.eigh(K)
    ev, U = ev[::-1][:k], U[:, ::-1][:, :k]
    loss = -ev.sum()
    Q = Wt.T @ (U / np.sqrt(ev))
    Y = Zc @ Q
    grads = []
    for i, s_ in enumerate(sl):
        A = (n - 1) * C[s_, :] @ Q
        B = Sinv[i] @ A
        g = -2.0 * Y @ B.T + (2.0 / (n - 1)) * Zc[:, s_] @ (B @ B.T)
        grads.append(g - g.mean(axis=0))
    return loss, grads


def ref_mccaloss(zs, eps=1e-5):
    """cca_zoo/deep/objectives.py:138-153."""
    tot = 0.0
    for i in range(len(zs)):
        for j in range(i + 1, len(zs)):
            tot += ref_ccaloss(zs[i], zs[j], eps)
    return tot


# --------------------------------------------------------------------------------------
# cov_* : covariance-space forms (what the CUDA kernels compute)
# --------------------------------------------------------------------------------------


def moments(views):
    """Raw block moments of the hstacked views: M = X^T X (D x D), s = 1^T X (D,), n.

    This is the quantity kernel K1 produces per row shard (and the quantity that is
    all-reduced across GPUs).  np.cov(hstack) of cca_zoo/linear/_mcca.py:150-151 and the
    per-view np.cov / SVDs elsewhere are all functions of it."""
    X = np.hstack([np.asarray(v, dtype=np.float64) for v in views])
    return X.T @ X, X.sum(axis=0), X.shape[0]


def covariance_from_moments(M, s, n, center=True):
    """C = (M - s s^T / n) / (n - 1)  (ddof=1, as np.cov / the 1/(n-1) in _rcca.py:96)."""
    if center:
        return (M - np.outer(s, s) / n) / (n - 1)
    return M / (n - 1)


def block_slices(dims):
    off = np.concatenate([[0], np.cumsum(dims)])
    return [slice(int(off[i]), int(off[i + 1])) for i in range(len(dims))]


def cov_whiten(Cii, c, n_samples, rank_tol=None):
    """Covariance form of svd_whiten (cca_zoo/_utils/_linalg.py:9-41):
    C = V diag(lam) V^T ; W = V diag(((1-c) lam + c)^-1/2), columns by DEscending lam
    (the SVD order).  The reference's ``s > 0`` filter (:30) only removes exact zeros in floating point; what makes
    a direction unusable is a numerically null REGULARISED eigenvalue, so directions are dropped when
    (1-c) lam + c <= tol ((1-c) lam_max + c): for c = 0 the plain relative filter (the reference itself returns
    1e13-sized weights there), for any practical ridge nothing is dropped -- as in the reference."""
    lam, V = np.linalg.eigh(Cii)
    lam, V = lam[::-1], V[:, ::-1]
    d = Cii.shape[0]
    if rank_tol is None:
        rank_tol = max(d, n_samples) * np.finfo(Cii.dtype).eps
    keep = (1.0 - c) * lam + c > rank_tol * ((1.0 - c) * max(lam[0], 0.0) + c)
    keep[min(d, n_samples):] = False  # thin SVD has at most min(n, d) directions
    lam, V = lam[keep], V[:, keep]
    g = 1.0 / np.sqrt((1.0 - c) * lam + c)
    return V * g, lam


def cov_rcca_fit(C, dims, latent_dimensions=1, c=0.0, n_samples=None):
    """Covariance form of rCCA.fit (cca_zoo/linear/_rcca.py:83-101)."""
    if len(dims) != 2:
        raise ValueError("rCCA requires exactly 2 views")
    c_ = perview(c, 0.0, 2)
    s1, s2 = block_slices(dims)
    n_samples = n_samples if n_samples is not None else 10**9
    W1, _ = cov_whiten(C[s1, s1], c_[0], n_samples)
    W2, _ = cov_whiten(C[s2, s2], c_[1], n_samples)
    k = min(latent_dimensions, W1.shape[1], W2.shape[1])
    T = W1.T @ C[s1, s2] @ W2
    U, sv, Vt = np.linalg.svd(T, full_matrices=False)
    return [W1 @ U[:, :k], W2 @ Vt[:k, :].T], sv[:k]


def cov_mcca_fit(C, dims, latent_dimensions=1, c=0.0, eps=1e-6):
    """Covariance form of MCCA.fit (cca_zoo/linear/_mcca.py:113-135,141-173):
    A = (C - blkdiag(C_ii))/m ; B = blkdiag((1-c_i) C_ii + c_i I)/m (+ eps floor) ;
    top-k of A v = lam B v with v^T B v = 1, descending."""
    m = len(dims)
    c_ = perview(c, 0.0, m)
    sl = block_slices(dims)
    D = C.shape[0]
    A = C.copy()
    B = np.zeros_like(C)
    for i, s in enumerate(sl):
        A[s, s] = 0.0
        B[s, s] = (1.0 - c_[i]) * C[s, s] + c_[i] * np.eye(dims[i])
    min_eig = min(np.linalg.eigvalsh(B[s, s]).min() for s in sl)
    if min_eig < eps:
        B = B + (eps - min_eig) * np.eye(D)
    A, B = A / m, B / m
    lam, vecs = ref_gevp(A, B, latent_dimensions)
    return [vecs[s, :] for s in sl], lam


def cov_gcca_fit(C, dims, n_samples, latent_dimensions=1, c=0.0, view_weights=None, eps=1e-6, second_moment=None):
    """Primal (D x D) restatement of GCCA.fit (cca_zoo/linear/_gcca.py:94-109;
    SURVEY.md §3.3).  R_i = ((1-c_i)C_ii + c_i I (+floor))^-1/2, S = blkdiag(sqrt(mu_i) R_i),
    G = (n-1) S C S ; top-k G U = U diag(sig) ; W_i = pinv(C_ii) [C S U]_i diag(sig)^-1/2.

    ``second_moment`` = X^T X/(n-1) without mean subtraction, for ``center=False``: the reference then still takes
    the regularised blocks from np.cov (centred, :98-100) but forms v R^-1 v^T and pinv(v) with the raw views
    (:105,109), so G, C S U and the pseudo-inverses use the second moment, S uses the covariance C."""
    Cd = C if second_moment is None else second_moment
    m = len(dims)
    c_ = perview(c, 0.0, m)
    mu = perview(view_weights, 1.0, m)
    sl = block_slices(dims)
    D = C.shape[0]
    S = np.zeros_like(C)
    pinvs = []
    for i, s in enumerate(sl):
        lam, V = np.linalg.eigh(C[s, s])
        reg = (1.0 - c_[i]) * lam + c_[i]
        if reg.min() < eps:
            reg = reg + (eps - reg.min())
        S[s, s] = np.sqrt(mu[i]) * (V / np.sqrt(reg)) @ V.T
        lam, V = np.linalg.eigh(Cd[s, s])
        tol = max(dims[i], n_samples) * np.finfo(C.dtype).eps * max(lam.max(), 0.0)
        inv = np.where(lam > tol, 1.0 / np.where(lam > tol, lam, 1.0), 0.0)
        pinvs.append((V * inv) @ V.T)
    G = (n_samples - 1) * (S @ Cd @ S)
    k = min(latent_dimensions, D, n_samples)
    sig, U = ref_gevp(G, None, k)
    CSU = Cd @ (S @ U)
    ws = [pinvs[i] @ CSU[s, :] / np.sqrt(sig) for i, s in enumerate(sl)]
    return ws, sig


def cov_ccaloss(z1, z2, eps=1e-5):
    """Closed form of CCALoss.forward valid whenever no clamp fires
    (cca_zoo/deep/objectives.py:20,101):
    loss = -tr(S11^-1 S12 S22^-1 S21).  Also returns the analytic gradients w.r.t. z1, z2."""
    z1 = np.asarray(z1, dtype=np.float64)
    z2 = np.asarray(z2, dtype=np.float64)
    n = z1.shape[0]
    a = z1 - z1.mean(axis=0)
    b = z2 - z2.mean(axis=0)
    s11 = a.T @ a / (n - 1) + eps * np.eye(a.shape[1])
    s22 = b.T @ b / (n - 1) + eps * np.eye(b.shape[1])
    s12 = a.T @ b / (n - 1)
    i11 = np.linalg.inv(s11)
    i22 = np.linalg.inv(s22)
    P = i11 @ s12 @ i22  # d1 x d2
    loss = -np.sum(P * s12)
    # dL/dS12 = -2 P ; dL/dS11 = P S21 S11^-1 (sym) ; dL/dS22 = S22^-1 S21 P (sym)
    g12 = -2.0 * P
    g11 = P @ s12.T @ i11
    g22 = i22 @ s12.T @ P
    # back through S = a^T b/(n-1) and the centring (projection onto 1-perp, a no-op on
    # already-centred cotangents because a, b are centred and the cotangents are linear in a, b)
    ga = (a @ (g11 + g11.T) + b @ g12.T) / (n - 1)
    gb = (b @ (g22 + g22.T) + a @ g12) / (n - 1)
    ga -= ga.mean(axis=0)
    gb -= gb.mean(axis=0)
    return loss, ga, gb


# --------------------------------------------------------------------------------------
# comparison helpers used by the parity tests
# --------------------------------------------------------------------------------------


def align_signs(weights, ref_weights):
    """Flip component signs jointly across views (a component flips in all views together)."""
    k = ref_weights[0].shape[1]
    out = [w.copy() for w in weights]
    for j in range(k):
        dot = sum(float(w[:, j] @ r[:, j]) for w, r in zip(weights, ref_weights))
        if dot < 0:
            for w in out:
                w[:, j] *= -1.0
    return out


def max_rel_err_per_vector(weights, ref_weights):
    ws = align_signs(weights, ref_weights)
    errs = []
    for w, r in zip(ws, ref_weights):
        errs.append(np.linalg.norm(w - r, axis=0) / np.maximum(np.linalg.norm(r, axis=0), 1e-300))
    return float(np.max(errs))


def subspace_distance(W, Wref):
    """|| P - Pref ||_2 between the column spans."""
    Q, _ = np.linalg.qr(W)
    Qr, _ = np.linalg.qr(Wref)
    return float(np.linalg.norm(Q @ Q.T - Qr @ Qr.T, 2))
