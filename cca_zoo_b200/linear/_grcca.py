"""Group-regularised CCA on the GPU (mirrors cca_zoo/linear/_grcca.py) -- SURVEY.md §8f: a caller of the MCCA core.

The reference augments every view with within-group residual columns and scaled group-mean columns, runs the
MCCA eigenproblem on the augmented views and folds the eigenvectors back.  Both steps are LINEAR in the features:
augmented_i = X_i T_i and weights_i = T_i block_i with

    T_i = [ (I - E diag(1/counts) E^T) / c_i  |  E diag(1 / sqrt(mu_i * counts)) ]      (d_i x (d_i + n_groups_i))

(E the feature -> group indicator; T_i = I when c_i <= 0).  So the augmented block covariance is
blkdiag(T)^T C blkdiag(T): no second pass over the samples, just 2 m GEMMs on the D x D covariance that the
moment kernel already produced.
"""
from __future__ import annotations

import warnings

import numpy as np
import torch

from .. import ops
from .._solvers import mcca_weights
from .._validation import perview_parameter
from ._mcca import MCCA


def group_map(n_features: int, groups, c: float, mu: float) -> np.ndarray:
    """T_i of the module docstring (float64, host: it depends on the labels only).
    Restates _augment_view / _collapse_weights (cca_zoo/linear/_grcca.py:126-160) as one matrix."""
    if c <= 0:
        return np.eye(n_features)
    labels = np.asarray(groups)
    if labels.shape != (n_features,):
        raise ValueError(f"feature_groups entry has shape {labels.shape}, expected ({n_features},)")
    _, inverse, counts = np.unique(labels, return_inverse=True, return_counts=True)
    E = np.zeros((n_features, counts.shape[0]))
    E[np.arange(n_features), inverse] = 1.0
    mu_eff = 1.0 if mu == 0 else float(mu)
    within = (np.eye(n_features) - (E / counts) @ E.T) / c
    between = E / np.sqrt(mu_eff * counts)
    return np.hstack([within, between])


class GRCCA(MCCA):
    """Group-regularised CCA (cca_zoo/linear/_grcca.py:18-176): ``fit(views, y=None, feature_groups=None)``;
    ``c`` is the within-group ridge, ``mu`` the group-mean penalty; ``weights_`` live in the original
    feature space."""

    # ``mu`` is not constrained in the reference either (it only declares MCCA's constraints)

    def __init__(self, latent_dimensions: int = 1, center: bool = True, c=0.0, mu=0.0, eps: float = 1e-6,
                 precision: str = "tf32x3b", device=None, solver: str = "auto") -> None:
        super().__init__(latent_dimensions=latent_dimensions, center=center, c=c, pca=False, eps=eps,
                         precision=precision, device=device, solver=solver)
        self.mu = mu

    def fit(self, views, y=None, feature_groups=None):
        C, dims, n_total = self._fit_device(views)
        m = len(dims)
        c_ = [float(x) for x in perview_parameter("c", self.c, 0.0, m)]
        mu_ = [float(x) for x in perview_parameter("mu", self.mu, 0.0, m)]
        if feature_groups is None:
            if any(ci > 0 for ci in c_):
                warnings.warn("No feature_groups provided; using a single group per view, which makes the "
                              "group regularisation a no-op.")
            feature_groups = [np.ones(d, dtype=int) for d in dims]
        if len(feature_groups) != m:
            raise ValueError(f"feature_groups must have one entry per view ({m}), got {len(feature_groups)}")
        self.feature_groups_ = feature_groups
        maps = [torch.from_numpy(group_map(d, g, ci, mi)).to(C.device, C.dtype)
                for d, g, ci, mi in zip(dims, feature_groups, c_, mu_)]
        adims = [int(T.shape[1]) for T in maps]
        off = np.concatenate([[0], np.cumsum(dims)]).astype(int)
        aoff = np.concatenate([[0], np.cumsum(adims)]).astype(int)
        D, Da = int(off[-1]), int(aoff[-1])
        if all(ci <= 0 for ci in c_):
            Ca = C
        else:
            Y = torch.empty((D, Da), dtype=C.dtype, device=C.device)            # C blkdiag(T)
            for j in range(m):
                ops.gemm(C[:, off[j]:off[j + 1]], maps[j], out=Y[:, aoff[j]:aoff[j + 1]])
            Ca = torch.empty((Da, Da), dtype=C.dtype, device=C.device)          # blkdiag(T)^T (C blkdiag(T))
            for i in range(m):
                ops.gemm(maps[i], Y[off[i]:off[i + 1]], transa=True, out=Ca[aoff[i]:aoff[i + 1]])
            Ca = 0.5 * (Ca + Ca.T)
        blocks = mcca_weights(Ca, adims, self.latent_dimensions, c_, float(self.eps), solver=self.solver)
        return self._finish([ops.gemm(T, b.contiguous()) for T, b in zip(maps, blocks)])

    def _solve(self, C, dims, n_total):  # partial_fit path: plain MCCA has no group labels to apply
        raise NotImplementedError("GRCCA.partial_fit is not supported: pass all rows to fit(views, feature_groups=...)")
