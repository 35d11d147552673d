"""GPU parity of PartialCCA / GRCCA (callers of the MCCA core) against the reference's golden
vectors (tests/golden/reference_outputs_ext.npz) and the oracle; the behavioural checks mirror the reference's
tests/linear/test_eigendecomposition.py:443-575.  Tolerances: 1e-5 (float64 inputs) / 1e-3 (float32 inputs)."""
import warnings

import numpy as np
import pytest

from oracle import restatement as R
from tests import golden_io as G

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(G.PARTIAL_CASES))
def test_partialcca_matches_reference_golden(name):
    from cca_zoo_b200.linear import PartialCCA

    case = G.PARTIAL_CASES[name]
    views, Z = G.ext_inputs(name)
    ref = G.ext_outputs(name)
    tol = 1e-3 if case["dtype"] == "f32" else 1e-5
    est = PartialCCA(**case["kwargs"]).fit(views, partials=Z)
    err = R.max_rel_err_per_vector([w.astype(np.float64) for w in est.weights_], ref["w"])
    assert err < tol, f"weights rel err {err:.2e}"
    for b, br in zip(est.confound_betas_, ref["beta"]):
        assert b.shape == br.shape
        np.testing.assert_allclose(b, br, rtol=tol, atol=tol * max(1e-2, np.abs(br).max()))
    for a, b in zip(est.means_, ref["mean"]):
        np.testing.assert_allclose(a, b, rtol=tol, atol=tol * 1e-2)
    zs = est.transform(views, partials=Z)
    pc = np.array([abs(np.corrcoef(zs[0][:, d], zs[1][:, d])[0, 1]) for d in range(zs[0].shape[1])])
    np.testing.assert_allclose(pc, ref["partial_corr"], rtol=tol)
    np.testing.assert_allclose(est.score(views), ref["score"], rtol=10 * tol, atol=tol)


@pytest.mark.parametrize("name", sorted(G.GROUP_CASES))
def test_grcca_matches_reference_golden(name):
    from cca_zoo_b200.linear import GRCCA

    case = G.GROUP_CASES[name]
    views, groups = G.ext_inputs(name)
    ref = G.ext_outputs(name)
    tol = 1e-3 if case["dtype"] == "f32" else 1e-5
    est = GRCCA(**case["kwargs"]).fit(views, feature_groups=groups)
    for w, wr in zip(est.weights_, ref["w"]):
        assert w.shape == wr.shape
    err = R.max_rel_err_per_vector([w.astype(np.float64) for w in est.weights_], ref["w"])
    assert err < tol, f"weights rel err {err:.2e}"
    np.testing.assert_allclose(est.score(views), ref["score"], rtol=tol)


def test_partialcca_reference_behaviour():
    from sklearn.utils._param_validation import InvalidParameterError

    from cca_zoo_b200.datasets import conftest_views
    from cca_zoo_b200.linear import PartialCCA

    two = conftest_views("two_views")
    Z = np.random.default_rng(1).standard_normal((50, 3))
    with pytest.raises(ValueError, match="partials"):
        PartialCCA(latent_dimensions=1).fit(two)
    with pytest.raises(InvalidParameterError):
        PartialCCA(c=2.0).fit(two, partials=Z)
    model = PartialCCA(latent_dimensions=2).fit(two, partials=Z)
    out = model.transform(two, partials=Z)
    assert [o.shape for o in out] == [(50, 2), (50, 2)]
    assert len(model.transform(two)) == 2                      # falls back to the plain projection
    ft = PartialCCA(latent_dimensions=2).fit_transform(two, partials=Z)
    for a, b in zip(ft, out):
        np.testing.assert_allclose(np.abs(a), np.abs(b), atol=1e-10)
    assert model.score(two).shape == (2,)
    # a 1-D confound vector is one column
    m1 = PartialCCA(latent_dimensions=1).fit(two, partials=Z[:, 0])
    m2 = PartialCCA(latent_dimensions=1).fit(two, partials=Z[:, :1])
    assert R.max_rel_err_per_vector(m1.weights_, m2.weights_) < 1e-12


def test_partialcca_removes_a_dominant_confound():
    from cca_zoo_b200.linear import PartialCCA

    rng = np.random.default_rng(0)
    n = 200
    z = rng.standard_normal((n, 2))
    confound = rng.standard_normal((n, 1))
    x1 = z @ rng.standard_normal((2, 6)) + confound @ rng.standard_normal((1, 6)) * 5.0 + 0.1 * rng.standard_normal((n, 6))
    x2 = z @ rng.standard_normal((2, 6)) + confound @ rng.standard_normal((1, 6)) * 5.0 + 0.1 * rng.standard_normal((n, 6))
    model = PartialCCA(latent_dimensions=2).fit([x1, x2], partials=confound)
    z1, z2 = model.transform([x1, x2], partials=confound)
    corrs = np.array([np.corrcoef(z1[:, d], z2[:, d])[0, 1] for d in range(2)])
    assert np.all(np.abs(corrs) > 0.5)
    w, _, _ = R.ref_partialcca_fit([x1, x2], confound, 2)
    assert R.max_rel_err_per_vector(model.weights_, w) < 1e-5


def test_partialcca_batches_tensors_and_rank_deficient_confounds():
    import torch

    from cca_zoo_b200.linear import PartialCCA

    rng = np.random.default_rng(11)
    n = 6000
    lat = rng.standard_normal((n, 3))
    Z = rng.standard_normal((n, 3)) + 0.5
    views = [lat @ rng.standard_normal((3, d)) + Z @ rng.standard_normal((3, d)) + rng.standard_normal((n, d))
             for d in (40, 24)]
    w_ref, _, b_ref = R.ref_partialcca_fit(views, Z, 3, 0.1)
    one = PartialCCA(latent_dimensions=3, c=0.1).fit(views, partials=Z)
    assert R.max_rel_err_per_vector(one.weights_, w_ref) < 1e-5
    inc = PartialCCA(latent_dimensions=3, c=0.1)
    for lo in range(0, n, 1500):
        inc.partial_fit([v[lo:lo + 1500] for v in views], partials=Z[lo:lo + 1500], solve=lo + 1500 >= n)
    assert R.max_rel_err_per_vector(inc.weights_, one.weights_) < 1e-8
    dev = PartialCCA(latent_dimensions=3, c=0.1).fit([torch.from_numpy(v).cuda() for v in views],
                                                     partials=torch.from_numpy(Z).cuda())
    assert R.max_rel_err_per_vector(dev.weights_, one.weights_) < 1e-10
    # a duplicated confound column: pinv semantics (minimum-norm betas), same residuals, same weights
    Zd = np.hstack([Z, Z[:, :1]])
    dup = PartialCCA(latent_dimensions=3, c=0.1).fit(views, partials=Zd)
    assert R.max_rel_err_per_vector(dup.weights_, w_ref) < 1e-5
    _, _, b_dup = R.ref_partialcca_fit(views, Zd, 3, 0.1)
    for a, b in zip(dup.confound_betas_, b_dup):
        np.testing.assert_allclose(a, b, atol=1e-6)


def test_partialcca_wide_views_use_the_cholesky_route():
    from cca_zoo_b200.linear import PartialCCA

    rng = np.random.default_rng(12)
    n, dims, q = 5000, (320, 256), 5
    lat = rng.standard_normal((n, 8))
    Z = rng.standard_normal((n, q)) + 1.0
    views = [((lat @ rng.standard_normal((8, d))) * 0.3 + Z @ rng.standard_normal((q, d)) + rng.standard_normal((n, d))
              ).astype(np.float32) for d in dims]
    w_ref, _, _ = R.ref_partialcca_fit([v.astype(np.float64) for v in views], Z, 4, 0.2)
    for solver in ("auto", "eigen"):
        est = PartialCCA(latent_dimensions=4, c=0.2, solver=solver).fit(views, partials=Z)
        err = R.max_rel_err_per_vector([w.astype(np.float64) for w in est.weights_], w_ref)
        assert err < 1e-3, f"{solver}: {err:.2e}"


def test_grcca_reference_behaviour():
    from cca_zoo_b200.datasets import conftest_views
    from cca_zoo_b200.linear import GRCCA, MCCA

    two, three = conftest_views("two_views"), conftest_views("three_views")
    rng = np.random.default_rng(2)
    g = [rng.integers(0, 3, size=v.shape[1]) for v in two]
    model = GRCCA(latent_dimensions=1, c=[0.5, 0.0]).fit(two, feature_groups=g)
    for w, v in zip(model.weights_, two):
        assert w.shape == (v.shape[1], 1)
    s_g = GRCCA(latent_dimensions=2, c=0.0).fit(two).score(two)
    s_m = MCCA(latent_dimensions=2, pca=False).fit(two).score(two)
    np.testing.assert_allclose(s_g, s_m, atol=1e-6)
    with pytest.warns(UserWarning, match="feature_groups"):
        GRCCA(latent_dimensions=1, c=0.5).fit(two)
    g3 = [np.random.default_rng(3).integers(0, 2, size=v.shape[1]) for v in three]
    assert len(GRCCA(latent_dimensions=1, c=0.3).fit(three, feature_groups=g3).transform(three)) == 3
    with pytest.raises(ValueError, match="feature_groups"):
        GRCCA(c=0.3).fit(two, feature_groups=g[:1])
    with pytest.raises(ValueError, match="shape"):
        GRCCA(c=0.3).fit(two, feature_groups=[g[0][:-1], g[1]])


def test_grcca_wide_views_against_oracle():
    from cca_zoo_b200.linear import GRCCA

    rng = np.random.default_rng(13)
    n, dims = 4000, (300, 260)
    lat = rng.standard_normal((n, 6))
    views = [(lat @ rng.standard_normal((6, d))) * 0.3 + rng.standard_normal((n, d)) for d in dims]
    groups = [rng.integers(0, 12, size=d) for d in dims]
    w_ref, _ = R.ref_grcca_fit(views, groups, 4, [0.3, 0.5], [1.0, 0.5])
    for solver in ("auto", "eigen"):
        with warnings.catch_warnings():
            warnings.simplefilter("error")
            est = GRCCA(latent_dimensions=4, c=[0.3, 0.5], mu=[1.0, 0.5], solver=solver).fit(views, feature_groups=groups)
        err = R.max_rel_err_per_vector(est.weights_, w_ref)
        assert err < 1e-5, f"{solver}: {err:.2e}"
