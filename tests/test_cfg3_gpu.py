"""BASELINE config 3 on the GPU: CCALoss / MCCALoss value and analytic gradient at batch 4096 for widths 64, 512,
two ragged pairs and a 3-view MCCALoss, against the float64 outputs of the reference's own forward + autograd
(tests/golden/reference_outputs_cfg3.npz).  Tolerances (north_star): 1e-5 for float64 inputs, 1e-3 for float32
inputs -- both against the FLOAT64 reference."""
import numpy as np
import pytest
import torch

from tests import golden_io as G

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(G.CFG3_CASES))
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-5), (torch.float32, 1e-3)])
def test_config3_loss_and_gradients(name, dtype, tol):
    from cca_zoo_b200.deep import CCALoss, MCCALoss

    c = G.CFG3_CASES[name]
    loss_ref, grads_ref = G.cfg3_outputs(name)
    zs = [z.to(dtype).cuda().requires_grad_(True) for z in G.cfg3_inputs(name)]
    fn = CCALoss(eps=c["eps"]) if c["kind"] == "cca" else MCCALoss(eps=c["eps"])
    loss = fn(zs)
    assert loss.dim() == 0 and loss.dtype == dtype
    assert abs(loss.item() - loss_ref) < tol * abs(loss_ref), f"loss {loss.item()} vs {loss_ref}"
    loss.backward()
    for i, (z, ref) in enumerate(zip(zs, grads_ref)):
        G.cfg3_check_gradient(z.grad.double().cpu().numpy(), ref, i, c, tol)


@pytest.mark.parametrize("widths", [(64, 64), (512, 512), (96, 160)])
def test_config3_upstream_gradient_and_repeatability(widths):
    """loss * 3 back-propagates 3 x the gradient; two evaluations of the same batch are bit-identical
    (fixed-order reductions everywhere on the path)."""
    from cca_zoo_b200.deep import CCALoss

    g = torch.Generator().manual_seed(7)
    lat = torch.randn(4096, 16, generator=g)
    zs = [(lat @ torch.randn(16, w, generator=g) + torch.randn(4096, w, generator=g)).cuda() for w in widths]
    outs = []
    for scale in (1.0, 3.0, 1.0):
        z = [t.clone().requires_grad_(True) for t in zs]
        loss = CCALoss()(z) * scale
        loss.backward()
        outs.append((loss.item(), z[0].grad.clone(), z[1].grad.clone()))
    assert outs[0][0] == outs[2][0]
    assert torch.equal(outs[0][1], outs[2][1]) and torch.equal(outs[0][2], outs[2][2])
    assert torch.allclose(outs[1][1], 3.0 * outs[0][1], rtol=1e-5, atol=1e-9)
