"""GPU parity of the estimators against the golden vectors made from the reference (tests/golden)
and against the oracle on fresh seeded inputs.  Tolerances follow BASELINE.json's north_star:
1e-5 relative for float64 inputs, 1e-3 for float32 inputs (weights per vector, sign-aligned jointly
across views, and canonical correlations)."""
import numpy as np
import pytest

from oracle import restatement as R
from tests import golden_io as G

pytestmark = pytest.mark.gpu


def _model(case, **extra):
    from cca_zoo_b200 import linear

    return getattr(linear, case["model"])(**case["kwargs"], **extra)


@pytest.mark.parametrize("name", sorted(G.CASES))
def test_fit_matches_reference_golden(name):
    case = G.CASES[name]
    views = G.case_inputs(name)
    w_ref, mu_ref, score_ref = G.case_outputs(name)
    est = _model(case).fit(views)
    tol = 1e-3 if case["dtype"] == "f32" else 1e-5
    assert len(est.weights_) == len(w_ref)
    for w, wr in zip(est.weights_, w_ref):
        assert w.shape == wr.shape
    err = R.max_rel_err_per_vector([w.astype(np.float64) for w in est.weights_], w_ref)
    assert err < tol, f"weights rel err {err:.2e}"
    np.testing.assert_allclose(est.score(views), score_ref, rtol=tol)
    for a, b in zip(est.means_, mu_ref):
        np.testing.assert_allclose(a, b, rtol=tol, atol=tol * 1e-2)
    assert est.n_samples_ == views[0].shape[0]
    assert est.n_features_in_ == [v.shape[1] for v in views]


@pytest.mark.parametrize("precision,tol", [("tf32x3", 1e-3), ("exact", 1e-3), ("tf32", 2e-2)])
def test_rcca_float32_precisions(precision, tol):
    """float32 views through each covariance arithmetic; single-pass TF32 is reported, not gated at 1e-3."""
    from cca_zoo_b200.linear import rCCA

    views = G.case_inputs("rcca_med32")
    w_ref, _, score_ref = G.case_outputs("rcca_med32")
    est = rCCA(latent_dimensions=6, c=0.1, precision=precision).fit(views)
    err = R.max_rel_err_per_vector([w.astype(np.float64) for w in est.weights_], w_ref)
    assert err < tol, f"{precision}: weights rel err {err:.2e}"
    np.testing.assert_allclose(est.score(views), score_ref, rtol=tol)
    assert est.weights_[0].dtype == np.float32  # the reference keeps float32 in rCCA


def test_quickstart_readme_scores():
    """BASELINE config 1 (README.md:52-72): train and held-out scores."""
    from cca_zoo_b200.linear import CCA

    views = G.dataset("quickstart")
    est = CCA(latent_dimensions=2).fit(views)
    np.testing.assert_allclose(est.score(views), [0.99551056, 0.99407941], atol=1e-7)
    test = [G.get("quickstart_test/v0"), G.get("quickstart_test/v1")]
    np.testing.assert_allclose(est.score(test), [0.97681356, 0.97297565], atol=1e-7)


def test_accepts_cuda_and_cpu_tensors():
    import torch
    from cca_zoo_b200.linear import rCCA

    views = G.dataset("two_views")
    ref = rCCA(latent_dimensions=2, c=0.1).fit(views)
    for conv in (lambda v: torch.from_numpy(v), lambda v: torch.from_numpy(v).cuda()):
        est = rCCA(latent_dimensions=2, c=0.1).fit([conv(v) for v in views])
        assert R.max_rel_err_per_vector(est.weights_, ref.weights_) < 1e-9


def test_medium_fresh_inputs_against_oracle_all_models():
    """Seeded inputs the fixtures do not contain: oracle (numpy) vs CUDA path, float64."""
    from cca_zoo_b200.datasets import joint_data
    from cca_zoo_b200.linear import GCCA, MCCA, rCCA

    views = joint_data(n_views=3, n_samples=4000, n_features=[150, 130, 70], latent_dimensions=8,
                       signal_to_noise=0.02, random_state=7)
    dims = [150, 130, 70]
    M, s, n = R.moments(views)
    C = R.covariance_from_moments(M, s, n)
    w, _ = R.cov_rcca_fit(C[:280, :280], dims[:2], 8, [0.2, 0.05], n)
    est = rCCA(latent_dimensions=8, c=[0.2, 0.05]).fit(views[:2])
    assert R.max_rel_err_per_vector(est.weights_, w) < 1e-5
    w, _ = R.cov_mcca_fit(C, dims, 8, 0.1)
    est = MCCA(latent_dimensions=8, c=0.1).fit(views)
    assert R.max_rel_err_per_vector(est.weights_, w) < 1e-5
    w, _ = R.cov_gcca_fit(C, dims, n, 8, 0.1, [1.0, 2.0, 0.5])
    est = GCCA(latent_dimensions=8, c=0.1, view_weights=[1.0, 2.0, 0.5]).fit(views)
    assert R.max_rel_err_per_vector(est.weights_, w) < 1e-5


def test_reference_property_checks():
    """The equivalences the reference's own tests pin (tests/linear/test_eigendecomposition.py)."""
    from cca_zoo_b200.linear import CCA, GCCA, MCCA, PLS, rCCA

    two = G.dataset("two_views")
    cca = CCA(latent_dimensions=2).fit(two)
    # rCCA(c=0) == CCA (:345-353); MCCA(2 views) == CCA (:411-416), atol 1e-6 on |scores|
    for other in (rCCA(latent_dimensions=2, c=0.0), MCCA(latent_dimensions=2)):
        o = other.fit(two)
        np.testing.assert_allclose(np.abs(o.score(two)), np.abs(cca.score(two)), atol=1e-6)
    # identical views -> correlation 1 (:330-336)
    ident = CCA(latent_dimensions=2).fit([two[0], two[0].copy()])
    np.testing.assert_allclose(ident.score([two[0], two[0]]), 1.0, atol=1e-6)
    # variates uncorrelated (:419-429)
    z = cca.transform(two)
    cc = np.corrcoef(z[0].T)
    assert abs(cc[0, 1]) < 1e-6
    # fit_transform == fit().transform() up to sign (:112-137)
    ft = CCA(latent_dimensions=2).fit_transform(two)
    np.testing.assert_allclose(np.abs(ft[0]), np.abs(z[0]), atol=1e-10)
    # 3 views rejected by 2-view models (:67-73)
    three = G.dataset("three_views")
    for cls in (CCA, rCCA, PLS):
        with pytest.raises(ValueError, match="exactly 2 views"):
            cls().fit(three)
    # GCCA == MCCA scores on three views (known answer)
    np.testing.assert_allclose(GCCA(latent_dimensions=2).fit(three).score(three),
                               MCCA(latent_dimensions=2).fit(three).score(three), atol=1e-8)


def test_latent_dimensions_clamped_like_reference():
    """k silently clamped to the smaller view width (_rcca.py:95, _linalg.py:65)."""
    from cca_zoo_b200.linear import MCCA, rCCA

    two = G.dataset("two_views")
    est = rCCA(latent_dimensions=20, c=0.1).fit(two)
    assert est.weights_[0].shape == (10, 8) and est.weights_[1].shape == (8, 8)
    est = MCCA(latent_dimensions=30).fit(two)
    assert est.weights_[0].shape == (10, 18)


@pytest.mark.parametrize("dtype,tol", [("f64", 1e-5), ("f32", 1e-3)])
def test_rcca_cholesky_route_equals_eigen_route(dtype, tol):
    """solver="cholesky" (Cholesky whitening + top-k subspace SVD) gives the weights of the eigen route
    (which is the one checked against the reference goldens) on a problem large enough to use it."""
    from cca_zoo_b200.datasets import joint_data
    from cca_zoo_b200.linear import rCCA

    views = joint_data(n_views=2, n_samples=6000, n_features=[320, 288], latent_dimensions=10,
                       signal_to_noise=0.02, random_state=11,
                       dtype=np.float32 if dtype == "f32" else np.float64)
    a = rCCA(latent_dimensions=10, c=0.1, solver="eigen").fit(views)
    b = rCCA(latent_dimensions=10, c=0.1, solver="cholesky").fit(views)
    auto = rCCA(latent_dimensions=10, c=0.1).fit(views)
    wa = [w.astype(np.float64) for w in a.weights_]
    assert R.max_rel_err_per_vector([w.astype(np.float64) for w in b.weights_], wa) < tol
    assert R.max_rel_err_per_vector([w.astype(np.float64) for w in auto.weights_], wa) < tol
    np.testing.assert_allclose(b.score(views), a.score(views), rtol=tol)
    if dtype == "f64":
        M, s, n = R.moments(views)
        w, _ = R.cov_rcca_fit(R.covariance_from_moments(M, s, n), [320, 288], 10, 0.1, n)
        assert R.max_rel_err_per_vector(b.weights_, w) < 1e-5


def test_rcca_cholesky_route_falls_back_when_rank_deficient():
    from cca_zoo_b200.linear import rCCA

    rng = np.random.default_rng(0)
    base = rng.standard_normal((900, 150))
    x1 = np.hstack([base, base @ rng.standard_normal((150, 150))])      # 300 columns of rank 150
    x2 = rng.standard_normal((900, 260)) + x1[:, :260]
    a = rCCA(latent_dimensions=3, c=0.0, solver="eigen").fit([x1, x2])
    b = rCCA(latent_dimensions=3, c=0.0, solver="cholesky").fit([x1, x2])
    np.testing.assert_allclose(b.score([x1, x2]), a.score([x1, x2]), rtol=1e-6)


def test_mcca_gcca_cholesky_route_equals_eigen_route_and_oracle():
    """solver="cholesky" (Cholesky reduction + top-k subspace iteration) vs solver="eigen" vs the oracle, float64."""
    from cca_zoo_b200.datasets import joint_data
    from cca_zoo_b200.linear import GCCA, MCCA

    dims = [220, 200, 180]
    views = joint_data(n_views=3, n_samples=5000, n_features=dims, latent_dimensions=6,
                       signal_to_noise=0.05, random_state=21)
    M, s, n = R.moments(views)
    C = R.covariance_from_moments(M, s, n)
    w_or, _ = R.cov_mcca_fit(C, dims, 6, [0.1, 0.0, 0.2])
    for solver in ("eigen", "cholesky", "auto"):
        est = MCCA(latent_dimensions=6, c=[0.1, 0.0, 0.2], solver=solver).fit(views)
        assert R.max_rel_err_per_vector(est.weights_, w_or) < 1e-5, solver
    w_or, _ = R.cov_gcca_fit(C, dims, n, 6, [0.1, 0.0, 0.2], [1.0, 2.0, 0.5])
    for solver in ("eigen", "cholesky", "auto"):
        est = GCCA(latent_dimensions=6, c=[0.1, 0.0, 0.2], view_weights=[1.0, 2.0, 0.5], solver=solver).fit(views)
        assert R.max_rel_err_per_vector(est.weights_, w_or) < 1e-5, solver


def test_mcca_cholesky_route_declines_when_eps_floor_may_be_active():
    """A nearly singular view: lambda_min(B) < eps, the reference adds the floor (_mcca.py:170-172); the
    Cholesky route cannot certify lambda_min and must hand over to the eigen route (same answer as eigen)."""
    from cca_zoo_b200.linear import MCCA

    rng = np.random.default_rng(5)
    base = rng.standard_normal((3000, 200))
    x1 = np.hstack([base, base[:, :100] + 1e-5 * rng.standard_normal((3000, 100))])   # 300 cols, 100 nearly dependent
    x2 = rng.standard_normal((3000, 280)) + np.hstack([base, base[:, :80]])
    a = MCCA(latent_dimensions=4, solver="eigen").fit([x1, x2])
    b = MCCA(latent_dimensions=4, solver="cholesky").fit([x1, x2])
    np.testing.assert_allclose(b.score([x1, x2]), a.score([x1, x2]), rtol=1e-8)


def test_edge_shapes():
    """Smallest legal inputs, single-column views, more columns than samples, too many views."""
    from cca_zoo_b200.linear import CCA, MCCA, rCCA

    rng = np.random.default_rng(0)
    # one column per view: the single canonical correlation is |Pearson r|
    x, y = rng.standard_normal((40, 1)), rng.standard_normal((40, 1))
    y = y + 0.8 * x
    est = CCA(latent_dimensions=1).fit([x, y])
    r = np.corrcoef(x[:, 0], y[:, 0])[0, 1]
    np.testing.assert_allclose(np.abs(est.score([x, y])), [abs(r)], atol=1e-10)
    # n < d with ridge: the covariance blocks are singular, the whitening must still be finite
    a, b = rng.standard_normal((20, 50)), rng.standard_normal((20, 30))
    est = rCCA(latent_dimensions=3, c=0.5).fit([a, b])
    assert all(np.isfinite(w).all() for w in est.weights_)
    from oracle import restatement as R
    w_ref, _ = R.ref_rcca_fit([a, b], 3, 0.5)
    assert R.max_rel_err_per_vector(est.weights_, w_ref) < 1e-6
    # minimum sample count for a covariance
    with pytest.raises(ValueError):
        CCA().fit([rng.standard_normal((1, 3)), rng.standard_normal((1, 2))])
    # more views than the moment kernel's tile table supports -> a clear error, not a wrong answer
    many = [rng.standard_normal((30, 2)) for _ in range(9)]
    with pytest.raises(ValueError, match="views"):
        MCCA().fit(many)
    # non-finite input in a CUDA tensor is rejected like check_array does for numpy
    import torch
    bad = torch.randn(30, 4, device="cuda")
    bad[3, 2] = float("nan")
    with pytest.raises(ValueError, match="NaN|infinity"):
        CCA().fit([bad, torch.randn(30, 3, device="cuda")])


def test_device_score_path_equals_reference_definition():
    """score() of CUDA inputs is computed from one moment pass; it must equal the
    sample-wise definition of cca_zoo/_base.py:153-194 (numpy path) for every estimator family."""
    import torch
    from cca_zoo_b200.linear import GCCA, MCCA, rCCA

    views = G.dataset("joint3_med")
    test = [v[:700] + 0.3 for v in views]                     # different rows, shifted means
    for est in (rCCA(latent_dimensions=4, c=0.1).fit(views[:2]), MCCA(latent_dimensions=3, c=0.05).fit(views),
                GCCA(latent_dimensions=3).fit(views)):
        vs = test[:2] if isinstance(est, rCCA) else test
        host = est.pairwise_correlations(vs)
        dev = est.pairwise_correlations([torch.from_numpy(v).cuda() for v in vs])
        np.testing.assert_allclose(dev, host, atol=1e-9)
        np.testing.assert_allclose(est.score([torch.from_numpy(v).cuda() for v in vs]), est.score(vs), atol=1e-9)


def test_partial_fit_and_streamed_fit_equal_one_shot_fit():
    """Row batches through partial_fit, and a host data set large enough to take the streamed (chunked H2D)
    path, give the weights of a single fit on device-resident data."""
    import torch
    from cca_zoo_b200.datasets import joint_data
    from cca_zoo_b200.linear import MCCA, rCCA

    views = joint_data(n_views=2, n_samples=9000, n_features=[64, 48], latent_dimensions=5,
                       signal_to_noise=0.1, random_state=4)
    ref = rCCA(latent_dimensions=5, c=0.1).fit([torch.from_numpy(v).cuda() for v in views])
    inc = rCCA(latent_dimensions=5, c=0.1)
    for lo, hi, last in [(0, 2500, False), (2500, 2501, False), (2501, 9000, True)]:
        inc.partial_fit([v[lo:hi] for v in views], solve=last)
    assert inc.n_samples_ == 9000
    assert R.max_rel_err_per_vector(inc.weights_, ref.weights_) < 1e-9
    np.testing.assert_allclose(inc.means_[0], ref.means_[0], rtol=1e-12, atol=1e-13)
    # another fit() forgets the partial state
    inc.fit([v[:100] for v in views])
    assert inc.n_samples_ == 100
    # streamed host path (threshold lowered so that the test stays small)
    st = rCCA(latent_dimensions=5, c=0.1)
    st._stream_threshold_bytes = 1 << 20
    st._stream_chunk_rows = 1000
    st.fit(views)
    assert R.max_rel_err_per_vector(st.weights_, ref.weights_) < 1e-9
    with pytest.raises(ValueError, match="exactly 2 views"):
        rCCA().partial_fit([views[0], views[1], views[0]])
    m = MCCA(latent_dimensions=3).partial_fit([v[:4000] for v in views]).partial_fit([v[4000:] for v in views])
    np.testing.assert_allclose(m.score(views), MCCA(latent_dimensions=3).fit(views).score(views), rtol=1e-9)


def test_fitted_and_partially_fitted_estimators_pickle():
    import pickle

    from cca_zoo_b200.linear import rCCA

    views = G.dataset("joint2_med")
    est = rCCA(latent_dimensions=3, c=0.1).fit(views)
    back = pickle.loads(pickle.dumps(est))
    np.testing.assert_array_equal(back.weights_[0], est.weights_[0])
    np.testing.assert_allclose(back.score(views), est.score(views))
    half = rCCA(latent_dimensions=3, c=0.1).partial_fit([v[:1500] for v in views], solve=False)
    resumed = pickle.loads(pickle.dumps(half)).partial_fit([v[1500:] for v in views])
    assert R.max_rel_err_per_vector(resumed.weights_, est.weights_) < 1e-9


def test_large_offsets_do_not_destroy_float32_parity():
    """float32 views whose means are 300 standard deviations away from zero.  The reference centres the data
    before anything else; the one-pass moment form needs the shifted accumulation to stay within the 1e-3 bar."""
    from cca_zoo_b200.linear import MCCA, rCCA

    rng = np.random.default_rng(12)
    z = rng.standard_normal((5000, 4)) * np.array([1.0, 0.8, 0.6, 0.45])
    views = [(z @ rng.standard_normal((4, d)) * 0.6 + rng.standard_normal((5000, d)) + 300.0 * (1 + np.arange(d) % 3))
             .astype(np.float32) for d in (40, 56, 24)]
    v64 = [v.astype(np.float64) for v in views]
    est = rCCA(latent_dimensions=3, c=0.05).fit(views[:2])
    w_ref, _ = R.ref_rcca_fit(v64[:2], 3, 0.05)
    assert R.max_rel_err_per_vector([w.astype(np.float64) for w in est.weights_], w_ref) < 1e-3
    m = MCCA(latent_dimensions=3, c=0.1).fit(views)
    w_ref, _ = R.ref_mcca_fit(v64, 3, 0.1)
    assert R.max_rel_err_per_vector(m.weights_, w_ref) < 1e-3


def test_transform_on_the_device_equals_the_reference_definition():
    """CUDA inputs (and large host inputs) are projected on the device -- one GEMM per view with the
    mean term folded in -- and equal (v - mean_) @ weights_ (cca_zoo/_base.py:108-123)."""
    import torch

    from cca_zoo_b200.linear import MCCA, rCCA

    views = G.dataset("joint3_med", "f32")
    for est in (rCCA(latent_dimensions=4, c=0.1).fit(views[:2]), MCCA(latent_dimensions=3, c=0.05).fit(views)):
        vs = views[:len(est.weights_)]
        host = est.transform(vs)                                        # numpy branch
        dev = est.transform([torch.from_numpy(v).cuda() for v in vs])   # device branch
        for h, d, v, m, w in zip(host, dev, vs, est.means_, est.weights_):
            ref = (v.astype(np.float64) - m.astype(np.float64)) @ w.astype(np.float64)
            assert isinstance(d, np.ndarray) and d.shape == h.shape and d.dtype == h.dtype
            scale = np.abs(ref).max()
            assert np.abs(d - ref).max() < 1e-4 * scale and np.abs(h - ref).max() < 1e-4 * scale


def test_gcca_ignores_a_view_with_zero_weight_like_the_reference():
    """view_weights with a zero entry -- the reference drops that view from the eigenproblem but still
    returns weights for it (cca_zoo/linear/_gcca.py:105,109); no division by zero on either solver route."""
    from cca_zoo_b200.linear import GCCA

    views = G.dataset("joint4_gcca", "f64")
    mu = [1.0, 0.0, 2.0, 1.0]
    w_ref, _ = R.ref_gcca_fit(views, 3, 0.1, mu)
    for solver in ("eigen", "cholesky"):
        est = GCCA(latent_dimensions=3, c=0.1, view_weights=mu, solver=solver).fit(views)
        assert all(np.isfinite(w).all() for w in est.weights_)
        assert R.max_rel_err_per_vector(est.weights_, w_ref) < 1e-6
