"""The fit behind the C ABI (ccab_rcca_fit, csrc/fit.cu): moments -> weights in one asynchronous library call.
Parity against the reference goldens / the float64 oracle, the status word, and a fit driven through ctypes only."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import restatement as R
from tests import golden_io as G

pytestmark = pytest.mark.gpu


def _views(n, dims, k=6, seed=0, dtype=np.float32):
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((n, k)) * np.linspace(1.0, 0.45, k)      # distinct signal strengths: separated correlations
    return [(z @ rng.standard_normal((k, d)) * 0.5 + rng.standard_normal((n, d)) + 0.3).astype(dtype) for d in dims]


@pytest.mark.parametrize("dtype,tol", [(np.float32, 1e-3), (np.float64, 1e-5)])
@pytest.mark.parametrize("dims,k,c", [([256, 256], 4, 0.1), ([384, 260], 5, 0.0), ([512, 300], 6, [0.3, 0.05])])
def test_device_fit_matches_the_oracle(dtype, tol, dims, k, c):
    from cca_zoo_b200.linear import rCCA

    views = _views(6000, dims, seed=len(dims) + k, dtype=dtype)
    est = rCCA(latent_dimensions=k, c=c).fit(views)
    assert est._fit_info["route"] == "device", est._fit_info
    w_ref, mu_ref = R.ref_rcca_fit([v.astype(np.float64) for v in views], k, c)
    assert R.max_rel_err_per_vector([w.astype(np.float64) for w in est.weights_], w_ref) < tol
    for m, mr in zip(est.means_, mu_ref):
        np.testing.assert_allclose(m, mr, rtol=1e-5, atol=1e-6)
    assert est.n_samples_ == 6000 and est.weights_[0].dtype == dtype and est.weights_[0].shape == (dims[0], k)
    sc = est.score(views)
    sc_ref = R.score(views, mu_ref, w_ref)
    np.testing.assert_allclose(sc, sc_ref, rtol=10 * tol)


def test_device_fit_declines_on_a_singular_block_and_the_host_route_takes_over():
    from cca_zoo_b200.linear import rCCA

    views = _views(3000, [260, 256], seed=3, dtype=np.float64)
    views[0][:, 7] = views[0][:, 3]                       # exactly singular covariance block, c = 0
    est = rCCA(latent_dimensions=4, c=0.0).fit(views)
    assert est._fit_info["route"] == "host" and est._fit_info["device_status"] & 1
    reduced = [np.delete(views[0], 7, axis=1), views[1]]
    ref = rCCA(latent_dimensions=4, c=0.0).fit(reduced)
    np.testing.assert_allclose(est.score(views), ref.score(reduced), rtol=1e-6)


def test_device_fit_reports_non_finite_input():
    from cca_zoo_b200.linear import rCCA

    views = [torch.from_numpy(v).cuda() for v in _views(2000, [256, 256], seed=5)]
    views[1][17, 3] = float("nan")
    with pytest.raises(ValueError, match="NaN or infinity"):
        rCCA(latent_dimensions=4, c=0.1).fit(views)


def test_fit_config1_through_ctypes_only():
    """BASELINE config 1 (README quickstart, CCA k=2 on 200 x [50, 50]) fitted by calling the C ABI directly: torch
    only owns the buffers.  This is what a non-Python host would do."""
    from cca_zoo_b200 import _lib

    lib = _lib.load()
    views = G.case_inputs("cca_quick")
    ws_ref, mus_ref, _ = G.case_outputs("cca_quick")
    dev = [torch.from_numpy(np.ascontiguousarray(v)).cuda() for v in views]
    n, dims = views[0].shape[0], [v.shape[1] for v in views]
    d = _lib.i64_array(dims)
    lds = _lib.i64_array([t.stride(0) for t in dev])
    ptrs = (C.c_void_p * 2)(*[t.data_ptr() for t in dev])
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    mom = torch.empty(lib.ccab_moments_size(2, d), dtype=torch.float64, device="cuda")
    ws = torch.empty(lib.ccab_moments_workspace_bytes(_lib.F64, _lib.PREC_EXACT, 2, d, n) + 256, dtype=torch.uint8,
                     device="cuda")
    assert lib.ccab_moments(_lib.F64, _lib.PREC_EXACT, 2, ptrs, d, lds, n, mom.data_ptr(), ws.data_ptr(), ws.numel(),
                            stream) == 0
    k, p = 2, 34
    offs = (C.c_int64 * 5)()
    assert lib.ccab_rcca_fit_result_layout(_lib.F64, d, k, p, offs) == 0
    block = torch.empty(offs[4] + 256, dtype=torch.uint8, device="cuda")
    block = block[(-block.data_ptr()) % 256:][:offs[4]]
    fws = torch.empty(lib.ccab_rcca_fit_workspace_bytes(_lib.F64, d, k, p) + 256, dtype=torch.uint8, device="cuda")
    cc = (C.c_double * 2)(0.0, 0.0)
    rc = lib.ccab_rcca_fit(_lib.F64, d, mom.data_ptr(), None, float(n), 1, cc, k, p, 16, block.data_ptr(), block.numel(),
                           fws.data_ptr(), fws.numel(), stream)
    assert rc == 0, _lib.last_error()
    host = block.cpu().numpy()
    hdr = host[:256].view(np.float64)
    assert int(hdr[0]) == 0 and int(hdr[1]) == n
    w = [host[offs[2 + i]:offs[2 + i] + 8 * dims[i] * k].view(np.float64).reshape(dims[i], k) for i in range(2)]
    mean = host[offs[0]:offs[0] + 8 * sum(dims)].view(np.float64)
    assert R.max_rel_err_per_vector(w, ws_ref) < 1e-6
    np.testing.assert_allclose(mean[:dims[0]], mus_ref[0], rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("dtype,tol", [(np.float32, 1e-3), (np.float64, 1e-5)])
@pytest.mark.parametrize("dims,k,c", [([256, 256, 256], 4, 0.1), ([320, 260, 200, 128], 5, [0.0, 0.05, 0.1, 0.2])])
def test_mcca_device_fit_matches_the_oracle(dtype, tol, dims, k, c):
    from cca_zoo_b200.linear import MCCA

    views = _views(7000, dims, seed=11 + len(dims), dtype=dtype)
    est = MCCA(latent_dimensions=k, c=c).fit(views)
    assert est._fit_info["route"] == "device", est._fit_info
    w_ref, mu_ref = R.ref_mcca_fit([v.astype(np.float64) for v in views], k, c)
    assert R.max_rel_err_per_vector([w.astype(np.float64) for w in est.weights_], w_ref) < tol
    assert est.weights_[0].dtype == np.float64          # np.cov upcasts in the reference: float64 weights
    np.testing.assert_allclose(est.score(views), R.score(views, mu_ref, w_ref), rtol=10 * tol)
    eig = MCCA(latent_dimensions=k, c=c, solver="eigen").fit(views)
    assert R.max_rel_err_per_vector(eig.weights_, w_ref) < tol


def test_loss_lazy_status_and_sync_mode_agree():
    from cca_zoo_b200.deep import CCALoss, MCCALoss

    g = torch.Generator().manual_seed(11)
    lat = torch.randn(2048, 8, generator=g)
    zs = [(lat @ torch.randn(8, w, generator=g) + torch.randn(2048, w, generator=g)).cuda() for w in (96, 72, 40)]
    out = {}
    for mode in ("lazy", "sync"):
        z = [t.clone().requires_grad_(True) for t in zs[:2]]
        fn = CCALoss(verify=mode)
        loss = fn(z)
        loss.backward()
        fn.check()
        out[mode] = (loss.item(), z[0].grad.clone())
    assert out["lazy"][0] == out["sync"][0] and torch.equal(out["lazy"][1], out["sync"][1])
    res = {}
    for mode in ("lazy", "sync"):                      # one moment pass + shared factorisations vs the pairwise loop
        z = [t.clone().requires_grad_(True) for t in zs]
        fn = MCCALoss(verify=mode)
        loss = fn(z)
        loss.backward()
        fn.check()
        res[mode] = (loss.item(), [t.grad.clone() for t in z])
    assert abs(res["lazy"][0] - res["sync"][0]) < 1e-4 * abs(res["sync"][0])
    for a, b in zip(res["lazy"][1], res["sync"][1]):
        assert float((a - b).abs().max() / b.abs().max()) < 1e-3


def test_loss_lazy_status_reports_a_broken_batch_at_the_next_call():
    from cca_zoo_b200.deep import CCALoss

    z1 = torch.randn(512, 16, device="cuda")
    z2 = torch.randn(512, 16, device="cuda")
    z2[5, 3] = float("inf")
    fn = CCALoss()
    fn([z1, z2])                                       # asynchronous: nothing is read back here
    with pytest.raises(ValueError, match="NaN or infinity"):
        fn.check()
