"""The reference's own behavioural tests for the estimators on the path, restated for this package and run on CPU
through the host logic (tests/fake_ops.py stands in for the kernels).  Sources: tests/test_sklearn_compat.py,
tests/linear/test_eigendecomposition.py:38-430, tests/linear/test_parameter_constraints.py of the reference.
The fixtures are the reference's conftest views (cca_zoo_b200.datasets.conftest_views reproduces them draw for draw).
"""
import numpy as np
import pytest
from sklearn.utils._param_validation import InvalidParameterError
from sklearn.utils.estimator_checks import (
    check_estimator_repr,
    check_get_params_invariance,
    check_no_attributes_set_in_init,
    check_set_params,
)

from cca_zoo_b200 import linear
from cca_zoo_b200.datasets import conftest_views
from tests import fake_ops

ALL = [getattr(linear, n) for n in linear.__all__]
TWO_VIEW = [linear.CCA, linear.rCCA, linear.PLS]
MULTI = [linear.MCCA, linear.GCCA]


@pytest.fixture
def host(monkeypatch):
    fake_ops.install(monkeypatch)


@pytest.fixture
def two_views():
    return conftest_views("two_views")


@pytest.fixture
def three_views():
    return conftest_views("three_views")


@pytest.fixture
def correlated_views():
    return conftest_views("correlated_views")


# ---- tests/test_sklearn_compat.py:59-75 -------------------------------------------------------------------------
@pytest.mark.parametrize("check", [check_no_attributes_set_in_init, check_get_params_invariance, check_set_params,
                                   check_estimator_repr], ids=lambda c: c.__name__)
@pytest.mark.parametrize("Model", ALL, ids=lambda c: c.__name__)
def test_sklearn_estimator_contract(Model, check):
    check(Model.__name__, Model())


# ---- tests/linear/test_parameter_constraints.py -------------------------------------------------------------------
@pytest.mark.parametrize("latent_dimensions", [0, -1, 1.5])
def test_invalid_latent_dimensions_rejected(host, latent_dimensions, two_views):
    with pytest.raises(InvalidParameterError):
        linear.CCA(latent_dimensions=latent_dimensions).fit(two_views)


def test_invalid_center_rejected(host, two_views):
    with pytest.raises(InvalidParameterError):
        linear.CCA(center="yes").fit(two_views)


@pytest.mark.parametrize("Model", [linear.rCCA, linear.MCCA, linear.GCCA, linear.GRCCA])
@pytest.mark.parametrize("c", [-0.1, 1.1])
def test_invalid_c_rejected(host, Model, c, two_views):
    with pytest.raises(InvalidParameterError):
        Model(c=c).fit(two_views)


@pytest.mark.parametrize("Model", [linear.MCCA, linear.GCCA])
def test_invalid_eps_rejected(host, Model, two_views):
    with pytest.raises(InvalidParameterError):
        Model(eps=0.0).fit(two_views)


def test_package_specific_parameters_are_validated_too(host, two_views):
    with pytest.raises(InvalidParameterError):
        linear.rCCA(precision="fp8").fit(two_views)
    with pytest.raises(InvalidParameterError):
        linear.MCCA(solver="lapack").fit(two_views)


# ---- tests/linear/test_eigendecomposition.py ----------------------------------------------------------------------
@pytest.mark.parametrize("Model", TWO_VIEW + MULTI)
def test_two_view_fit_transform_score_shapes(host, Model, two_views):
    k = 2
    model = Model(latent_dimensions=k).fit(two_views)
    out = model.transform(two_views)
    assert [o.shape for o in out] == [(50, k), (50, k)]
    assert [w.shape for w in model.weights_] == [(10, k), (8, k)]
    score = model.score(two_views)
    assert score.shape == (k,) and np.all(score >= -1 - 1e-9) and np.all(score <= 1 + 1e-9)
    assert model.pairwise_correlations(two_views).shape == (2, 2, k)
    loadings = model.get_factor_loadings(two_views)
    assert [ld.shape for ld in loadings] == [(10, k), (8, k)]
    for a, b in zip(Model(latent_dimensions=k).fit_transform(two_views), out):
        np.testing.assert_allclose(np.abs(a), np.abs(b), atol=1e-8)


@pytest.mark.parametrize("Model", MULTI)
def test_three_view_shapes(host, Model, three_views):
    model = Model(latent_dimensions=2).fit(three_views)
    assert [o.shape for o in model.transform(three_views)] == [(50, 2)] * 3
    assert model.score(three_views).shape == (2,)
    assert model.pairwise_correlations(three_views).shape == (3, 3, 2)
    assert [w.shape for w in model.weights_] == [(10, 2), (8, 2), (6, 2)]


@pytest.mark.parametrize("Model", TWO_VIEW)
def test_two_view_models_reject_three_views(host, Model, three_views):
    with pytest.raises(ValueError, match="exactly 2 views"):
        Model().fit(three_views)


@pytest.mark.parametrize("k", [1, 3, 5])
def test_multiple_latent_dimensions(host, k, two_views):
    assert linear.CCA(latent_dimensions=k).fit(two_views).transform(two_views)[0].shape == (50, k)


@pytest.mark.parametrize("c", [0.0, 0.5, 1.0])
def test_rcca_c_parameter(host, c, two_views):
    assert linear.rCCA(latent_dimensions=2, c=c).fit(two_views).score(two_views).shape == (2,)


def test_rcca_per_view_c_and_wrong_length(host, two_views):
    linear.rCCA(latent_dimensions=1, c=[0.1, 0.9]).fit(two_views)
    with pytest.raises(ValueError, match="length 2"):
        linear.rCCA(c=[0.1, 0.2, 0.3]).fit(two_views)


@pytest.mark.parametrize("pca", [True, False])
def test_mcca_pca_flag_gives_the_same_scores(host, pca, two_views):
    a = linear.MCCA(latent_dimensions=2, pca=pca).fit(two_views).score(two_views)
    b = linear.MCCA(latent_dimensions=2, pca=not pca).fit(two_views).score(two_views)
    np.testing.assert_allclose(a, b, atol=1e-10)


def test_gcca_view_weights(host, three_views):
    model = linear.GCCA(latent_dimensions=1, view_weights=[1.0, 2.0, 0.5]).fit(three_views)
    assert len(model.transform(three_views)) == 3


def test_cca_on_correlated_and_identical_views(host, correlated_views):
    score = linear.CCA(latent_dimensions=2).fit(correlated_views).score(correlated_views)
    assert np.all(score > 0.9) and score[0] >= score[1] - 1e-12
    x = np.random.default_rng(0).standard_normal((100, 5))
    np.testing.assert_allclose(linear.CCA(latent_dimensions=1).fit([x, x.copy()]).score([x, x.copy()]), [1.0],
                               atol=1e-6)


def test_rcca_zero_regularisation_matches_cca_and_mcca_two_views_matches_cca(host, correlated_views):
    s_cca = linear.CCA(latent_dimensions=2).fit(correlated_views).score(correlated_views)
    s_r = linear.rCCA(latent_dimensions=2, c=0.0).fit(correlated_views).score(correlated_views)
    s_m = linear.MCCA(latent_dimensions=2).fit(correlated_views).score(correlated_views)
    np.testing.assert_allclose(s_r, s_cca, atol=1e-10)
    np.testing.assert_allclose(s_m, s_cca, atol=1e-6)


@pytest.mark.parametrize("Model", TWO_VIEW + MULTI)
def test_center_false(host, Model, two_views):
    model = Model(latent_dimensions=1, center=False).fit(two_views)
    assert all(np.all(m == 0) for m in model.means_)
    assert model.transform(two_views)[0].shape == (50, 1)


def test_cca_canonical_variates_are_uncorrelated(host):
    rng = np.random.default_rng(42)
    z = rng.standard_normal((200, 3))
    x1 = z @ rng.standard_normal((3, 8)) + 0.5 * rng.standard_normal((200, 8))
    x2 = z @ rng.standard_normal((3, 6)) + 0.5 * rng.standard_normal((200, 6))
    z1, _ = linear.CCA(latent_dimensions=3).fit([x1, x2]).transform([x1, x2])
    corr = np.corrcoef(z1.T)
    assert np.abs(corr - np.diag(np.diag(corr))).max() < 1e-8
