"""GPU parity of CCALoss / MCCALoss (value and analytic gradient) against the float64 reference outputs
stored in tests/golden (made by oracle/make_golden.py with the reference's autograd)."""
import numpy as np
import pytest
import torch

from tests import golden_io as G

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(G.LOSS_CASES))
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-5), (torch.float32, 1e-3)])
def test_loss_and_gradients(name, dtype, tol):
    from cca_zoo_b200.deep import CCALoss, MCCALoss

    c = G.LOSS_CASES[name]
    loss_ref, grads_ref = G.loss_outputs(name)
    zs = [z.to(dtype).cuda().requires_grad_(True) for z in G.loss_inputs(name)]
    fn = CCALoss(eps=c["eps"]) if c["kind"] == "cca" else MCCALoss(eps=c["eps"])
    loss = fn(zs)
    assert loss.dim() == 0
    assert abs(loss.item() - loss_ref) < tol * abs(loss_ref)
    loss.backward()
    for z, gr in zip(zs, grads_ref):
        g = z.grad.double().cpu().numpy()
        denom = np.abs(gr).max()
        assert np.abs(g - gr).max() < tol * denom, f"grad err {np.abs(g - gr).max() / denom:.2e}"


def test_loss_is_nonpositive_scalar_and_rejects_wrong_view_count():
    """tests/deep/test_deep.py:214-239 of the reference."""
    from cca_zoo_b200.deep import CCALoss

    z = [torch.randn(40, 5, device="cuda") for _ in range(3)]
    loss = CCALoss()(z[:2])
    assert loss.dim() == 0 and loss.item() <= 0
    with pytest.raises(ValueError, match="exactly 2"):
        CCALoss()(z)


def test_loss_known_answer():
    """CCALoss(eps=1e-4) on torch.manual_seed(0) randn(16,4) float64."""
    from cca_zoo_b200.deep import CCALoss

    torch.manual_seed(0)
    z1 = torch.randn(16, 4, dtype=torch.float64)
    z2 = torch.randn(16, 4, dtype=torch.float64)
    loss = CCALoss(eps=1e-4)([z1.cuda(), z2.cuda()])
    assert abs(loss.item() - (-0.8498609668152676)) < 1e-10


def test_loss_trains_an_encoder():
    """The objective plugs into a plain torch training loop (the DCCA seam, cca_zoo/deep/_dcca.py:73-92)."""
    from cca_zoo_b200.deep import CCALoss

    torch.manual_seed(0)
    zl = torch.randn(512, 3, device="cuda")
    x1 = zl @ torch.randn(3, 10, device="cuda") + 0.1 * torch.randn(512, 10, device="cuda")
    x2 = zl @ torch.randn(3, 12, device="cuda") + 0.1 * torch.randn(512, 12, device="cuda")
    e1, e2 = torch.nn.Linear(10, 3).cuda(), torch.nn.Linear(12, 3).cuda()
    opt = torch.optim.Adam(list(e1.parameters()) + list(e2.parameters()), lr=1e-2)
    fn = CCALoss()
    first = None
    for _ in range(60):
        opt.zero_grad()
        loss = fn([e1(x1), e2(x2)])
        loss.backward()
        opt.step()
        first = loss.item() if first is None else first
    assert loss.item() < first - 0.1


def test_loss_rank_deficient_batch_uses_eigen_route_and_matches_reference_forward():
    """Fewer samples than features: S is singular, the eigenvalue clamp of the reference is active and the
    Cholesky shortcut must not be taken.  Forward value against the oracle's literal restatement."""
    from cca_zoo_b200.deep import CCALoss
    from oracle import restatement as R

    g = torch.Generator().manual_seed(4)
    z1 = torch.randn(6, 9, generator=g, dtype=torch.float64)
    z2 = torch.randn(6, 7, generator=g, dtype=torch.float64)
    loss = CCALoss(eps=1e-3)([z1.cuda(), z2.cuda()])
    ref = R.ref_ccaloss(z1.numpy(), z2.numpy(), 1e-3)
    assert abs(loss.item() - ref) < 1e-6 * abs(ref)


@pytest.mark.parametrize("name", sorted(G.GLOSS_CASES))
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-5), (torch.float32, 1e-3)])
def test_gccaloss_and_gradients(name, dtype, tol):
    """GCCALoss (primal D x D form, analytic backward) against the reference's n x n forward and autograd."""
    from cca_zoo_b200.deep import GCCALoss

    c = G.GLOSS_CASES[name]
    loss_ref, grads_ref = G.loss_outputs(name)
    zs = [z.to(dtype).cuda().requires_grad_(True) for z in G.loss_inputs(name)]
    loss = GCCALoss(eps=c["eps"])(zs)
    assert loss.dim() == 0 and loss.dtype == dtype
    assert abs(loss.item() - loss_ref) < tol * abs(loss_ref)
    loss.backward()
    for z, gr in zip(zs, grads_ref):
        g = z.grad.double().cpu().numpy()
        denom = np.abs(gr).max()
        assert np.abs(g - gr).max() < tol * denom, f"grad err {np.abs(g - gr).max() / denom:.2e}"


def test_gccaloss_reference_behaviour_and_large_batch():
    """tests/deep/test_deep.py:252-258 of the reference (scalar for 3 views) + a batch size whose n x n form
    (the reference's) would need 134 MB and an O(n^3) eigensolve per step; checked against the oracle's primal
    form in float64."""
    from cca_zoo_b200.deep import GCCALoss
    from oracle import restatement as R

    views = [torch.randn(32, 4, device="cuda") for _ in range(3)]
    loss = GCCALoss(eps=1e-4)(views)
    assert loss.dim() == 0 and loss.item() < 0
    g = torch.Generator().manual_seed(5)
    lat = torch.randn(4096, 6, generator=g, dtype=torch.float64)
    zs = [lat @ torch.randn(6, w, generator=g, dtype=torch.float64)
          + 0.7 * torch.randn(4096, w, generator=g, dtype=torch.float64) for w in (48, 64, 40)]
    L, grads = R.cov_gccaloss([z.numpy() for z in zs], 1e-5)
    zc = [z.cuda().requires_grad_(True) for z in zs]
    loss = GCCALoss()(zc)
    loss.backward()
    assert abs(loss.item() - L) < 1e-8 * abs(L)
    for z, gr in zip(zc, grads):
        assert np.abs(z.grad.cpu().numpy() - gr).max() < 1e-7 * np.abs(gr).max()
    with pytest.raises(RuntimeError, match="CUDA"):
        GCCALoss()([z.cpu() for z in views])
