"""Oracle restatements of PartialCCA / GRCCA (callers of the MCCA core) against the golden
vectors made from the reference by oracle/make_golden_ext.py.  Runs everywhere (no GPU, no /root/reference)."""
import numpy as np
import pytest

from oracle import restatement as R
from tests import golden_io as G


def _tol(case):
    # float32 inputs: the reference itself works in float64 after np.cov / pinv, the inputs are exact in both
    return 1e-6 if case["dtype"] == "f32" else 1e-8


@pytest.mark.parametrize("form", ["ref", "cov"])
@pytest.mark.parametrize("name", sorted(G.PARTIAL_CASES))
def test_partialcca_oracle_matches_golden(name, form):
    case = G.PARTIAL_CASES[name]
    views, Z = G.ext_inputs(name)
    ref = G.ext_outputs(name)
    kw = dict(case["kwargs"])
    k, center, c = kw.pop("latent_dimensions"), kw.pop("center", True), kw.pop("c", 0.0)
    v64 = [v.astype(np.float64) for v in views]
    if form == "ref":
        w, _, betas = R.ref_partialcca_fit(v64, Z, k, c, center=center)
    else:
        M, s, n = R.moments(v64 + [Z])
        w, betas = R.cov_partialcca(M, s, n, [v.shape[1] for v in views], Z.shape[1], k, c, center=center)
    assert R.max_rel_err_per_vector(w, ref["w"]) < _tol(case)
    for a, b in zip(betas, ref["beta"]):
        np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-6 if case["dtype"] == "f32" else 1e-10)


@pytest.mark.parametrize("form", ["ref", "cov"])
@pytest.mark.parametrize("name", sorted(G.GROUP_CASES))
def test_grcca_oracle_matches_golden(name, form):
    case = G.GROUP_CASES[name]
    views, groups = G.ext_inputs(name)
    ref = G.ext_outputs(name)
    kw = dict(case["kwargs"])
    k, c, mu = kw.pop("latent_dimensions"), kw.pop("c", 0.0), kw.pop("mu", 0.0)
    v64 = [v.astype(np.float64) for v in views]
    if form == "ref":
        w, means = R.ref_grcca_fit(v64, groups, k, c, mu)
    else:
        M, s, n = R.moments(v64)
        C = R.covariance_from_moments(M, s, n, True)
        w = R.cov_grcca(C, [v.shape[1] for v in views], groups, k, c, mu)
        means = [v.mean(axis=0) for v in v64]
    assert R.max_rel_err_per_vector(w, ref["w"]) < _tol(case)
    np.testing.assert_allclose(R.score(v64, means, w), ref["score"], rtol=1e-6)


def test_grcca_zero_c_is_plain_mcca():
    """tests/linear/test_eigendecomposition.py:553-558 of the reference, on the oracle."""
    views, groups = G.ext_inputs("grcca_c0")
    w, _ = R.ref_grcca_fit(views, groups, 2, 0.0)
    w_mcca, _ = R.ref_mcca_fit(views, 2, 0.0)
    assert R.max_rel_err_per_vector(w, w_mcca) < 1e-12


def test_partialcca_removes_a_dominant_confound():
    """tests/linear/test_eigendecomposition.py:493-516 of the reference, on the oracle (same seeded data)."""
    rng = np.random.default_rng(0)
    n = 200
    z = rng.standard_normal((n, 2))
    confound = rng.standard_normal((n, 1))
    x1 = z @ rng.standard_normal((2, 6)) + confound @ rng.standard_normal((1, 6)) * 5.0 + 0.1 * rng.standard_normal((n, 6))
    x2 = z @ rng.standard_normal((2, 6)) + confound @ rng.standard_normal((1, 6)) * 5.0 + 0.1 * rng.standard_normal((n, 6))
    M, s, nn = R.moments([x1, x2, confound])
    w, betas = R.cov_partialcca(M, s, nn, [6, 6], 1, 2)
    mu = [x1.mean(0), x2.mean(0)]
    z1, z2 = [((x - m) - confound @ b) @ ww for x, m, b, ww in zip([x1, x2], mu, betas, w)]
    corrs = np.array([abs(np.corrcoef(z1[:, d], z2[:, d])[0, 1]) for d in range(2)])
    assert np.all(corrs > 0.5)


@pytest.mark.parametrize("name", sorted(G.GLOSS_CASES))
def test_gccaloss_oracle_matches_golden(name):
    """Literal n x n restatement and the primal form (with its analytic gradient) against the reference's
    forward value and autograd gradients."""
    c = G.GLOSS_CASES[name]
    zs = [z.numpy() for z in G.loss_inputs(name)]
    loss_ref, grads_ref = G.loss_outputs(name)
    assert abs(R.ref_gccaloss(zs, c["eps"]) - loss_ref) < 1e-10 * abs(loss_ref)
    L, grads = R.cov_gccaloss(zs, c["eps"])
    assert abs(L - loss_ref) < 1e-10 * abs(loss_ref)
    for g, gr in zip(grads, grads_ref):
        np.testing.assert_allclose(g, gr, rtol=1e-6, atol=1e-9 * np.abs(gr).max())


def _center_case_weights(case, views, form):
    """center=False semantics per estimator (see oracle/make_golden_ext.py: CENTER_CASES)."""
    kw = dict(case["kwargs"])
    k, center, c = kw.pop("latent_dimensions"), kw.pop("center", True), kw.pop("c", 0.0)
    model = case["model"]
    v64 = [v.astype(np.float64) for v in views]
    dims = [v.shape[1] for v in views]
    if form == "ref":
        if model == "rCCA":
            return R.ref_rcca_fit(v64, k, c, center=center)[0]
        if model == "MCCA":
            return R.ref_mcca_fit(v64, k, c, kw.get("eps", 1e-6), center=center)[0]
        return R.ref_gcca_fit(v64, k, c, kw.get("view_weights"), kw.get("eps", 1e-6), center=center)[0]
    M, s, n = R.moments(v64)
    C = R.covariance_from_moments(M, s, n, True)
    Cu = R.covariance_from_moments(M, s, n, False)
    if model == "rCCA":      # tall SVD of the views as they are: second moments when center=False
        return R.cov_rcca_fit(C if center else Cu, dims, k, c, n)[0]
    if model == "MCCA":      # np.cov centres whatever `center` says
        return R.cov_mcca_fit(C, dims, k, c, kw.get("eps", 1e-6))[0]
    return R.cov_gcca_fit(C, dims, n, k, c, kw.get("view_weights"), kw.get("eps", 1e-6),
                          second_moment=None if center else Cu)[0]


@pytest.mark.parametrize("form", ["ref", "cov"])
@pytest.mark.parametrize("name", sorted(G.CENTER_CASES))
def test_center_false_and_ridge_rank_deficiency_oracle_matches_golden(name, form):
    case = G.CENTER_CASES[name]
    views, _ = G.ext_inputs(name)
    ref = G.ext_outputs(name)
    w = _center_case_weights(case, views, form)
    assert [x.shape for x in w] == [x.shape for x in ref["w"]]
    # rcca_dup_ridge: the 9th singular value of the whitened cross-covariance is exactly zero and its left vector is
    # arbitrary in a 2-dimensional null space -- compare the 8 determined directions
    kk = 8 if name == "rcca_dup_ridge" else w[0].shape[1]
    tol = 1e-5 if case["dtype"] == "f32" else 1e-7
    assert R.max_rel_err_per_vector([x[:, :kk] for x in w], [x[:, :kk] for x in ref["w"]]) < tol
