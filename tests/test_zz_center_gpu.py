"""GPU parity of the ``center=False`` semantics (np.cov centres inside MCCA / GCCA / GRCCA whatever ``center`` says;
GCCA additionally mixes in raw second moments) and of the ridge-regularised fit of a rank-deficient view, against
goldens made from the reference (oracle/make_golden_ext.py: CENTER_CASES).  Named to sort last: these semantics
were added after the round's last GPU session and validated on the torch-CPU stand-in only
(tools/run_gpu_tests_on_standin.py); the kernels they reach are the ones every other test exercises."""
import numpy as np
import pytest

from oracle import restatement as R
from tests import golden_io as G

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(G.CENTER_CASES))
def test_center_semantics_match_reference_golden(name):
    from cca_zoo_b200 import linear

    case = G.CENTER_CASES[name]
    views, _ = G.ext_inputs(name)
    ref = G.ext_outputs(name)
    tol = 1e-3 if case["dtype"] == "f32" else 1e-5
    est = getattr(linear, case["model"])(**case["kwargs"]).fit(views)
    assert [w.shape for w in est.weights_] == [w.shape for w in ref["w"]]
    # rcca_dup_ridge: the 9th singular value is exactly zero, its left vector arbitrary in a 2-d null space
    kk = 8 if name == "rcca_dup_ridge" else est.weights_[0].shape[1]
    err = R.max_rel_err_per_vector([w[:, :kk].astype(np.float64) for w in est.weights_],
                                   [w[:, :kk] for w in ref["w"]])
    assert err < tol, f"weights rel err {err:.2e}"
    np.testing.assert_allclose(est.score(views)[:kk], ref["score"][:kk], rtol=10 * tol, atol=tol)
    for a, b in zip(est.means_, ref["mean"]):
        np.testing.assert_allclose(a, b, rtol=tol, atol=tol * 1e-2)


def test_center_false_solver_routes_agree():
    """Wide views so that ``auto`` takes the Cholesky route: same weights as the eigen route and as the oracle."""
    from cca_zoo_b200.linear import GCCA, MCCA

    rng = np.random.default_rng(21)
    n, dims = 3000, (280, 260, 300)
    z = rng.standard_normal((n, 5))
    views = [z @ rng.standard_normal((5, d)) * 0.4 + rng.standard_normal((n, d)) + 0.7 for d in dims]
    w_m, _ = R.ref_mcca_fit(views, 3, 0.1, center=False)
    w_g, _ = R.ref_gcca_fit(views, 3, 0.1, center=False)
    for solver in ("cholesky", "eigen"):
        m = MCCA(latent_dimensions=3, c=0.1, center=False, solver=solver).fit(views)
        g = GCCA(latent_dimensions=3, c=0.1, center=False, solver=solver).fit(views)
        assert R.max_rel_err_per_vector(m.weights_, w_m) < 1e-5, solver
        assert R.max_rel_err_per_vector(g.weights_, w_g) < 1e-5, solver
        assert all(np.all(mu == 0) for mu in m.means_ + g.means_)
