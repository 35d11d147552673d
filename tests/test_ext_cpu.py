"""Host-side logic of the PartialCCA / GRCCA wrappers that needs no GPU."""
import numpy as np
import pytest

from oracle import restatement as R


def test_group_map_equals_oracle_maps_and_reference_augmentation():
    from cca_zoo_b200.linear._grcca import group_map

    rng = np.random.default_rng(0)
    d = 11
    g = rng.integers(0, 4, size=d)
    for c, mu in [(0.4, 0.0), (0.7, 2.5), (0.0, 1.0)]:
        T = group_map(d, g, c, mu)
        np.testing.assert_allclose(T, R.grcca_maps([d], [g], [c], [mu])[0], atol=1e-15)
    # the augmentation written the way the reference writes it (per-group means), not through the oracle's map
    X = rng.standard_normal((30, d))
    ids, inv, counts = np.unique(g, return_inverse=True, return_counts=True)
    gm = np.stack([X[:, g == i].mean(axis=1) for i in ids], axis=1)
    aug = np.hstack([(X - gm[:, inv]) / 0.4, gm / np.sqrt(1.0 / counts)])
    np.testing.assert_allclose(X @ group_map(d, g, 0.4, 0.0), aug, atol=1e-12)


def test_partialcca_argument_errors_need_no_device():
    from cca_zoo_b200.linear import PartialCCA

    v = [np.zeros((6, 3)), np.zeros((6, 2))]
    with pytest.raises(ValueError, match="partials"):
        PartialCCA().fit(v)
    with pytest.raises(ValueError, match="rows"):
        PartialCCA().fit(v, partials=np.zeros((5, 1)))
    with pytest.raises(ValueError, match="at most 7 views"):
        PartialCCA().fit([np.zeros((6, 2))] * 8, partials=np.zeros((6, 1)))


def test_new_estimators_keep_the_sklearn_contract():
    from sklearn.base import clone

    from cca_zoo_b200.linear import GRCCA, PartialCCA

    p = clone(PartialCCA(latent_dimensions=3, c=[0.1, 0.2]))
    assert p.get_params()["c"] == [0.1, 0.2] and "pca" not in p.get_params()
    g = clone(GRCCA(c=0.2, mu=[1.0, 2.0])).set_params(mu=0.5)
    assert g.get_params()["mu"] == 0.5 and g.pca is False
