"""GCCA on the GPU (mirrors cca_zoo/linear/_gcca.py)."""
from __future__ import annotations

from typing import Any, ClassVar

from sklearn.utils._param_validation import StrOptions

from .._base import BaseModel
from .._solvers import gcca_weights
from .._validation import perview_parameter
from ._mcca import POSITIVE_EPS
from ._rcca import RIDGE_PARAMETER


class GCCA(BaseModel):
    r"""Generalised (MAX-VAR) CCA (cca_zoo/linear/_gcca.py:16-110).

    The reference builds the :math:`n \times n` matrix :math:`Q=\sum_i \mu_i X_i
    ((1-c_i)X_i^\top X_i + c_i I)^{-1} X_i^\top` and back-solves with ``pinv``; here the identical
    weights come from the :math:`D \times D` primal form, so ``n_samples`` is no
    longer bounded by an :math:`O(n^2)` allocation.  Eigen-stage and ``weights_`` are float64.
    """

    _solve_in_float64 = True
    _covariance_always_centred = True      # np.cov for the regularised blocks (cca_zoo/linear/_gcca.py:98-100)
    _wants_second_moment = True            # ... but raw products for Q and pinv when center=False (:105,109)
    _parameter_constraints: ClassVar[dict[str, list[Any]]] = {
        **BaseModel._parameter_constraints,
        "c": RIDGE_PARAMETER,
        "view_weights": [None, "array-like"],
        "eps": POSITIVE_EPS,
        "solver": [StrOptions({"auto", "eigen", "cholesky"})],
    }

    def __init__(self, latent_dimensions: int = 1, center: bool = True, c=0.0, view_weights=None,
                 eps: float = 1e-6, precision: str = "tf32x3b", device=None, solver: str = "auto") -> None:
        super().__init__(latent_dimensions=latent_dimensions, center=center, precision=precision, device=device)
        self.c = c
        self.view_weights = view_weights
        self.eps = eps
        self.solver = solver

    def fit(self, views, y=None):
        C, dims, n_total = self._fit_device(views)
        return self._finish(self._solve(C, dims, n_total))

    def _solve(self, C, dims, n_total):
        c_ = perview_parameter("c", self.c, 0.0, self.n_views_)
        mu = perview_parameter("view_weights", self.view_weights, 1.0, self.n_views_)
        second, self._second_moment = getattr(self, "_second_moment", None), None
        return gcca_weights(C, dims, n_total, self.latent_dimensions, [float(x) for x in c_],
                            [float(x) for x in mu], float(self.eps), solver=self.solver, second_moment=second)
