"""Partial CCA on the GPU (mirrors cca_zoo/linear/_partialcca.py) -- SURVEY.md §8f: a caller of the MCCA core.

The reference regresses the confounds P out of every centred view (``pinv(P) @ Xc``, a tall least-squares per
view) and runs the MCCA eigenproblem on the residuals.  In covariance space no residual is ever formed: P joins
the views as one more column group of the block-moment pass (K1), and with

    G    = P^T Xc                      (q x D, from the cross block of the moments)
    beta = (P^T P)^+ G                 (== pinv(P) @ Xc, the reference's ``confound_betas_``)
    t    = 1^T (Xc - P beta)           (column sums of the residuals; np.cov centres them again)

the covariance of the residuals is  (Xc^T Xc - G^T beta - t t^T / n) / (n - 1)  -- a q x q eigenproblem and two
small GEMMs on top of the same single pass over the data, additive over row shards like everything else here.
"""
from __future__ import annotations

import numpy as np
import torch
from sklearn.utils.validation import check_is_fitted

from .. import ops, parallel
from .._solvers import mcca_weights
from .._validation import perview_parameter, validate_views
from ._mcca import MCCA


def _pinv_psd_apply(Mpp: torch.Tensor, G: torch.Tensor) -> torch.Tensor:
    """(P^T P)^+ G through the Jacobi eigensolver (K3); directions below q * eps * lambda_max are dropped (numpy's
    pinv cuts singular values of P at 1e-15 * sigma_max, which no Gram-matrix method can resolve)."""
    q = Mpp.shape[0]
    if q == 1:
        val = Mpp[0, 0]
        return torch.where(val > 0, G / val, torch.zeros_like(G))
    lam, vt = ops.syevj(Mpp)
    tol = q * torch.finfo(Mpp.dtype).eps * lam[0]
    inv = torch.where(lam > tol, 1.0 / lam.clamp_min(torch.finfo(Mpp.dtype).tiny), torch.zeros_like(lam))
    Z = ops.gemm(vt, G)                                    # q x D : V^T G
    Z = ops.scale(Z, rows=inv)
    return ops.gemm(vt, Z, transa=True)                    # V diag(1/lam) V^T G


class PartialCCA(MCCA):
    """CCA of the views after removing the linear effect of confounds (cca_zoo/linear/_partialcca.py:17-160).

    Same constructor as the reference (``latent_dimensions, center, c, eps``; ``pca`` is fixed to False there)
    plus this package's ``precision / device / solver``.  ``fit`` needs ``partials`` (n_samples x n_confounds);
    ``transform(views, partials=None)`` deconfounds only when partials are given, like the reference.
    Under ``torch.distributed`` the views AND the partials are this rank's row shard.
    """

    def __init__(self, latent_dimensions: int = 1, center: bool = True, c=0.0, eps: float = 1e-6,
                 precision: str = "tf32x3b", device=None, solver: str = "auto") -> None:
        super().__init__(latent_dimensions=latent_dimensions, center=center, c=c, pca=False, eps=eps,
                         precision=precision, device=device, solver=solver)

    # ------------------------------------------------------------------ fit
    def _validated_with_partials(self, views, partials):
        if partials is None:
            raise ValueError("PartialCCA requires `partials` to be provided to fit().")
        validated = validate_views(views)
        if len(validated) + 1 > 8:
            raise ValueError("PartialCCA supports at most 7 views (the moment kernel takes 8 column groups).")
        if isinstance(partials, torch.Tensor):
            P = partials if partials.dim() == 2 else partials.reshape(partials.shape[0], -1)
            if not P.dtype.is_floating_point:
                P = P.to(torch.float64)
        else:
            P = np.asarray(partials, dtype=float)
            if P.ndim == 1:
                P = P[:, None]
            if P.ndim != 2:
                raise ValueError(f"partials must be 2-D (n_samples, n_confounds), got shape {P.shape}")
        if P.shape[0] != validated[0].shape[0]:
            raise ValueError(f"partials have {P.shape[0]} rows, the views {validated[0].shape[0]}")
        # one dtype for the whole moment pass: float32 views keep the tensor-core path, the confounds follow
        first = validated[0]
        f32 = (first.dtype == torch.float32) if isinstance(first, torch.Tensor) else (first.dtype == np.float32)
        if f32 and all((v.dtype in (torch.float32, np.float32)) for v in validated):
            P = P.to(torch.float32) if isinstance(P, torch.Tensor) else P.astype(np.float32)
        return validated, P

    def fit(self, views, y=None, partials=None):
        self._validate_params()
        validated, P = self._validated_with_partials(views, partials)
        device = self._device()
        mom, n_local, dims_all, in_dtype = self._local_moments(validated + [P], device)
        self._partial = None
        return self._solve_partial(mom, n_local, dims_all, in_dtype)

    def partial_fit(self, views, y=None, partials=None, solve: bool = True):
        """Row batches of (views, partials); see BaseModel.partial_fit."""
        self._validate_params()
        validated, P = self._validated_with_partials(views, partials)
        device = self._device()
        mom, n_local, dims_all, in_dtype = self._local_moments(validated + [P], device)
        state = getattr(self, "_partial", None)
        if state is not None:
            if state["dims"] != dims_all or state["dtype"] != in_dtype:
                raise ValueError(f"partial_fit batches must keep the widths/dtype: {state['dims']} vs {dims_all}")
            mom = state["mom"].to(device).add_(mom)
            n_local += state["n"]
        self._partial = {"mom": mom, "n": n_local, "dims": dims_all, "dtype": in_dtype}
        if solve:
            self._solve_partial(mom.clone(), n_local, dims_all, in_dtype)
        return self

    def _solve_partial(self, mom, n_local, dims_all, in_dtype):
        mom, n_total = parallel.allreduce_moments(mom, n_local)
        if not bool(torch.isfinite(mom).all()):
            raise ValueError("Input contains NaN or infinity.")
        dims, q = dims_all[:-1], dims_all[-1]
        D = int(sum(dims))
        f64 = torch.float64
        Cc, mean = ops.covariance(mom, dims_all, n_total, center=True, dtype=f64)
        n, nm1 = float(n_total), float(n_total - 1)
        pbar = mean[D:]
        if self.center:
            Cxx = Cc[:D, :D]
            G = Cc[D:, :D] * nm1                                        # P^T Xc = Pc^T Xc
            Mpp = Cc[D:, D:] * nm1 + n * torch.outer(pbar, pbar)         # P^T P (uncentred)
        else:
            Cu, _ = ops.covariance(mom, dims_all, n_total, center=False, dtype=f64)
            Cxx = Cu[:D, :D]
            G = Cu[D:, :D] * nm1
            Mpp = Cu[D:, D:] * nm1
        G = G.contiguous()
        beta = _pinv_psd_apply(Mpp.contiguous(), G)                      # q x D
        colsum = torch.zeros(D, dtype=f64, device=mom.device) if self.center else mean[:D] * n
        t = colsum - n * ops.gemm(pbar[None, :].contiguous(), beta)[0]
        Cd = Cxx.clone()
        ops.gemm(G, beta, transa=True, alpha=-1.0 / nm1, beta=1.0, out=Cd)
        Cd.sub_(torch.outer(t, t) / (n * nm1))
        Cd = 0.5 * (Cd + Cd.T)

        self.n_views_ = len(dims)
        self.n_features_in_ = list(dims)
        self.n_samples_ = n_total
        off = np.concatenate([[0], np.cumsum(dims)]).astype(int)
        np_dtype = np.float32 if in_dtype == torch.float32 else np.float64
        mean_np = mean[:D].cpu().numpy()
        if self.center:
            self.means_ = [mean_np[off[i]:off[i + 1]].astype(np_dtype) for i in range(len(dims))]
        else:
            self.means_ = [np.zeros(p) for p in dims]
        beta_np = beta.cpu().numpy()
        self.confound_betas_ = [beta_np[:, off[i]:off[i + 1]] for i in range(len(dims))]
        c_ = perview_parameter("c", self.c, 0.0, self.n_views_)
        weights = mcca_weights(Cd, list(dims), self.latent_dimensions, [float(x) for x in c_], float(self.eps),
                               solver=self.solver)
        return self._finish(weights)

    # ------------------------------------------------------------------ transform (reference semantics)
    def transform(self, views, partials=None):
        check_is_fitted(self)
        if partials is None:
            return super().transform(views)
        validated = validate_views(self._as_numpy_views(views))
        if isinstance(partials, torch.Tensor):
            partials = partials.detach().cpu().numpy()
        P = np.asarray(partials, dtype=float)
        if P.ndim == 1:
            P = P[:, None]
        out = []
        for v, mu, b, w in zip(validated, self.means_, self.confound_betas_, self.weights_):
            out.append(((v - mu) - P @ b) @ w)
        return out

    def fit_transform(self, views, y=None, partials=None):
        return self.fit(views, y=y, partials=partials).transform(views, partials=partials)
