"""The oracle (oracle/restatement.py) against the golden vectors made from the reference.

Runs everywhere (no GPU, no /root/reference).  Tolerances: float64 LAPACK round-off for the
f64 cases; the f32 cases compare the oracle's float64 covariance-space result with the
reference's own float32 run, i.e. they measure the reference's float32 noise (<= 1e-3).
"""
import numpy as np
import pytest

from oracle import restatement as R
from tests import golden_io as G

LINEAR = sorted(G.CASES)


def _fit_oracle(case, views, form):
    kw = dict(case["kwargs"])
    k = kw.pop("latent_dimensions")
    center = kw.pop("center", True)
    model = case["model"]
    if model == "CCA":
        kw["c"] = 0.0
    if model == "PLS":
        kw["c"] = 1.0
    kw.pop("pca", None)
    dims = [v.shape[1] for v in views]
    n = views[0].shape[0]
    if form == "ref":
        v64 = views if model in ("CCA", "rCCA", "PLS") else [v.astype(np.float64) for v in views]
        fn = {"CCA": R.ref_rcca_fit, "rCCA": R.ref_rcca_fit, "PLS": R.ref_rcca_fit,
              "MCCA": R.ref_mcca_fit, "GCCA": R.ref_gcca_fit}[model]
        return fn(v64, k, center=center, **kw)
    M, s, n = R.moments(views)
    C = R.covariance_from_moments(M, s, n, center)
    means = [v.mean(axis=0) if center else np.zeros(v.shape[1]) for v in views]
    if model in ("CCA", "rCCA", "PLS"):
        w, _ = R.cov_rcca_fit(C, dims, k, kw.get("c", 0.0), n)
    elif model == "MCCA":
        w, _ = R.cov_mcca_fit(C, dims, k, kw.get("c", 0.0), kw.get("eps", 1e-6))
    else:
        w, _ = R.cov_gcca_fit(C, dims, n, k, kw.get("c", 0.0), kw.get("view_weights"),
                              kw.get("eps", 1e-6))
    return w, means


@pytest.mark.parametrize("form", ["ref", "cov"])
@pytest.mark.parametrize("name", LINEAR)
def test_linear_oracle_matches_golden(name, form):
    case = G.CASES[name]
    views = G.case_inputs(name)
    w_ref, mu_ref, score_ref = G.case_outputs(name)
    w, mu = _fit_oracle(case, views, form)
    f32 = case["dtype"] == "f32" and case["model"] in ("CCA", "rCCA", "PLS")
    # MCCA/GCCA upcast to float64 in np.cov, but the reference's PCA pre-rotation of a
    # float32 view runs in float32 (sklearn keeps the dtype): ~1e-6 noise in its weights.
    tol = 2e-3 if f32 else (1e-5 if case["dtype"] == "f32" else 1e-8)
    assert R.max_rel_err_per_vector(w, w_ref) < tol
    for a, b in zip(mu, mu_ref):
        in32 = case["dtype"] == "f32"  # float32 column means carry float32 round-off
        np.testing.assert_allclose(a, b, rtol=1e-5 if in32 else 1e-12, atol=2e-6 if in32 else 1e-12)
    sc = R.score(views, mu, w)
    np.testing.assert_allclose(sc, score_ref, rtol=1e-4 if f32 else (1e-6 if case["dtype"] == "f32" else 1e-9))


def test_quickstart_heldout_score():
    """README.md:52-72 -- score on a fresh sample of the same generator."""
    views = G.dataset("quickstart")
    w, mu = R.ref_rcca_fit(views, 2, 0.0)
    test = [G.get("quickstart_test/v0"), G.get("quickstart_test/v1")]
    sc = R.score(test, mu, w)
    np.testing.assert_allclose(sc, G.get("quickstart_test/score"), rtol=1e-9)
    np.testing.assert_allclose(sc, [0.97681356, 0.97297565], atol=1e-8)


@pytest.mark.parametrize("name", sorted(G.LOSS_CASES))
def test_loss_oracle_matches_golden(name):
    c = G.LOSS_CASES[name]
    zs = [z.numpy() for z in G.loss_inputs(name)]
    loss_ref, grads_ref = G.loss_outputs(name)
    if c["kind"] == "cca":
        assert abs(R.ref_ccaloss(zs[0], zs[1], c["eps"]) - loss_ref) < 1e-10 * abs(loss_ref)
        L, ga, gb = R.cov_ccaloss(zs[0], zs[1], c["eps"])
        assert abs(L - loss_ref) < 1e-10 * abs(loss_ref)
        np.testing.assert_allclose(ga, grads_ref[0], rtol=1e-7, atol=1e-12)
        np.testing.assert_allclose(gb, grads_ref[1], rtol=1e-7, atol=1e-12)
    else:
        assert abs(R.ref_mccaloss(zs, c["eps"]) - loss_ref) < 1e-10 * abs(loss_ref)
        grads = [np.zeros_like(z) for z in zs]
        tot = 0.0
        for i in range(len(zs)):
            for j in range(i + 1, len(zs)):
                L, ga, gb = R.cov_ccaloss(zs[i], zs[j], c["eps"])
                tot += L
                grads[i] += ga
                grads[j] += gb
        assert abs(tot - loss_ref) < 1e-10 * abs(loss_ref)
        for g, gr in zip(grads, grads_ref):
            np.testing.assert_allclose(g, gr, rtol=1e-7, atol=1e-12)


def test_torch_restatement_of_the_loss_matches_reference_autograd():
    """oracle.ref_ccaloss_torch_fwdbwd (the CPU arm of bench.py's config-3 workloads) against the reference's own
    forward + autograd gradients stored in tests/golden."""
    for name, c in G.LOSS_CASES.items():
        if c["kind"] != "cca":
            continue
        loss_ref, grads_ref = G.loss_outputs(name)
        z = G.loss_inputs(name)
        L, ga, gb = R.ref_ccaloss_torch_fwdbwd(z[0], z[1], c["eps"])
        assert abs(float(L) - loss_ref) < 1e-10 * abs(loss_ref)
        np.testing.assert_allclose(ga.numpy(), grads_ref[0], rtol=1e-7, atol=1e-11)
        np.testing.assert_allclose(gb.numpy(), grads_ref[1], rtol=1e-7, atol=1e-11)


def test_known_answers_from_survey():
    """Known answers recorded in SURVEY.md §8c from the reference run."""
    from cca_zoo_b200.datasets import conftest_views

    cv = conftest_views("correlated_views")
    w, mu = R.ref_rcca_fit(cv, 2, 0.0)
    np.testing.assert_allclose(R.score(cv, mu, w), [0.9993354815, 0.9975869168], atol=1e-9)
    w, mu = R.ref_rcca_fit(cv, 2, 0.1)
    np.testing.assert_allclose(R.score(cv, mu, w), [0.9991982166, 0.9967209599], atol=1e-9)
    w, mu = R.ref_rcca_fit(cv, 2, 1.0)
    np.testing.assert_allclose(R.score(cv, mu, w), [0.9799367365, 0.9775105998], atol=1e-9)
    tv = conftest_views("three_views")
    w, mu = R.ref_mcca_fit(tv, 2, 0.0)
    np.testing.assert_allclose(R.score(tv, mu, w), [0.5467570014, 0.4299724741], atol=1e-9)
    w, mu = R.ref_gcca_fit(tv, 2, 0.0)
    np.testing.assert_allclose(R.score(tv, mu, w), [0.5467570014, 0.4299724741], atol=1e-9)
    import torch

    torch.manual_seed(0)
    z1 = torch.randn(16, 4, dtype=torch.float64).numpy()
    z2 = torch.randn(16, 4, dtype=torch.float64).numpy()
    assert abs(R.ref_ccaloss(z1, z2, 1e-4) - (-0.8498609668152676)) < 1e-12
