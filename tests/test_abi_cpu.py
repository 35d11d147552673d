"""The C ABI without a GPU: the shared library loads, exports every symbol include/ccab200.h declares,
answers the pure host queries, and every compute entry point fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

from cca_zoo_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ccab200.h")).read()
    return sorted(set(re.findall(r"\b(ccab_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ccab200.h but not exported"
    assert set(_lib.SIGNATURES) == set(declared), "ctypes table and header disagree"


def test_host_queries():
    lib = _lib.load()
    assert lib.ccab_version() >= 100
    assert lib.ccab_moments_padded_dim(2, _lib.i64_array([1024, 1024])) == 2048
    assert lib.ccab_moments_padded_dim(3, _lib.i64_array([10, 8, 6])) == 384
    assert lib.ccab_moments_size(2, _lib.i64_array([50, 50])) == 256 * 256 + 256
    assert lib.ccab_moments_size(9, _lib.i64_array([4] * 9)) == -1  # more than CCAB_MAX_VIEWS
    assert "n_views" in _lib.last_error()
    assert lib.ccab_moments_workspace_bytes(0, 0, 2, _lib.i64_array([1024, 1024]), 100000) > 0
    assert lib.ccab_syevj_workspace_bytes(0, 1024, 2) > 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_calls_fail_loudly_without_gpu():
    lib = _lib.load()
    buf = (C.c_double * 16)()
    rc = lib.ccab_gemm(_lib.F64, 0, 0, 2, 2, 2, 1.0, C.cast(buf, C.c_void_p), 2, C.cast(buf, C.c_void_p), 2, 0.0,
                       C.cast(buf, C.c_void_p), 2, None)
    assert rc != 0
    assert "CUDA" in _lib.last_error() or "device" in _lib.last_error()
    from cca_zoo_b200 import ops
    from cca_zoo_b200.deep import CCALoss
    from cca_zoo_b200.linear import rCCA

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.moments([torch.zeros(4, 4), torch.zeros(4, 4)])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rCCA().fit([torch.zeros(8, 3).numpy(), torch.zeros(8, 2).numpy()])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        CCALoss()([torch.zeros(8, 3), torch.zeros(8, 3)])


def test_estimator_surface_matches_reference_contract():
    """sklearn contract the reference tests pin (tests/test_sklearn_compat.py:59-75, test_base.py:49-114)."""
    import numpy as np
    from sklearn.base import clone
    from sklearn.exceptions import NotFittedError
    from sklearn.utils._param_validation import InvalidParameterError

    from cca_zoo_b200.linear import CCA, GCCA, MCCA, PLS, rCCA

    for cls in (CCA, rCCA, PLS, MCCA, GCCA):
        est = cls(latent_dimensions=3)
        assert clone(est).get_params() == est.get_params()
        assert "latent_dimensions=3" in repr(est)
        est.set_params(latent_dimensions=2)
        assert est.latent_dimensions == 2
        assert not any(k.endswith("_") for k in vars(est)), "no fitted attributes in __init__"
        with pytest.raises(NotFittedError):
            est.transform([np.zeros((4, 3)), np.zeros((4, 2))])
    two = [np.random.default_rng(0).standard_normal((20, 4)), np.random.default_rng(1).standard_normal((20, 3))]
    # parameter constraints are enforced at fit time, before any device work
    with pytest.raises(InvalidParameterError):
        rCCA(latent_dimensions=0).fit(two)
    with pytest.raises(InvalidParameterError):
        rCCA(c=1.5).fit(two)
    with pytest.raises(InvalidParameterError):
        MCCA(eps=0.0).fit(two)
    # view validation messages of cca_zoo/_utils/_validation.py:31-41
    if not torch.cuda.is_available():
        with pytest.raises((ValueError, RuntimeError)):
            rCCA().fit([two[0]])
    from cca_zoo_b200._validation import perview_parameter, validate_views

    with pytest.raises(ValueError, match="At least 2 views"):
        validate_views([two[0]])
    with pytest.raises(ValueError, match="same number of samples"):
        validate_views([two[0], two[1][:10]])
    with pytest.raises(ValueError, match="NaN|infinity|inf"):
        validate_views([two[0], np.full((20, 3), np.nan)])
    assert perview_parameter("c", 0.5, 0.0, 3) == [0.5, 0.5, 0.5]
    assert perview_parameter("c", None, 0.0, 2) == [0.0, 0.0]
    with pytest.raises(ValueError, match="list of length"):
        perview_parameter("c", [0.1], 0.0, 2)


def test_estimators_pickle_roundtrip():
    """Fitted state is plain numpy (picklable like the reference's sklearn estimators)."""
    import pickle

    import numpy as np

    from cca_zoo_b200.linear import rCCA

    est = rCCA(latent_dimensions=2, c=0.1)
    est.weights_ = [np.ones((4, 2)), np.ones((3, 2))]
    est.means_ = [np.zeros(4), np.zeros(3)]
    est.n_views_, est.n_features_in_, est.n_samples_ = 2, [4, 3], 10
    back = pickle.loads(pickle.dumps(est))
    assert back.get_params() == est.get_params()
    two = [np.random.default_rng(0).standard_normal((5, 4)), np.random.default_rng(1).standard_normal((5, 3))]
    np.testing.assert_allclose(back.transform(two)[0], est.transform(two)[0])
