"""The oracle's covariance-form objective (cov_ccaloss: value + analytic gradient, the form the CUDA kernels
implement) against the reference's own forward / autograd at BASELINE config 3 sizes
(tests/golden/reference_outputs_cfg3.npz, made by oracle/make_golden_cfg3.py from /root/reference)."""
import numpy as np
import pytest

from oracle import restatement as R
from tests import golden_io as G


@pytest.mark.parametrize("name", sorted(G.CFG3_CASES))
def test_oracle_matches_reference_at_config3(name):
    c = G.CFG3_CASES[name]
    loss_ref, grads_ref = G.cfg3_outputs(name)
    zs = [z.numpy() for z in G.cfg3_inputs(name)]
    if c["kind"] == "cca":
        L, ga, gb = R.cov_ccaloss(zs[0], zs[1], c["eps"])
        grads = [ga, gb]
    else:
        grads = [np.zeros_like(z) for z in zs]
        L = 0.0
        for i in range(len(zs)):
            for j in range(i + 1, len(zs)):
                Lij, ga, gb = R.cov_ccaloss(zs[i], zs[j], c["eps"])
                L += Lij
                grads[i] += ga
                grads[j] += gb
    assert abs(L - loss_ref) < 1e-10 * abs(loss_ref)
    for i, (g, ref) in enumerate(zip(grads, grads_ref)):
        G.cfg3_check_gradient(g, ref, i, c, 1e-8)
