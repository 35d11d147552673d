"""Does single-pass TF32 miss the 1e-3 weight tolerance at config-2 scale because of the ROUNDING MODE of the operands?

The tensor core truncates fp32 operands to TF32 (tools/probe_trunc.py).  Truncation is biased: every product is
under-estimated by ~2^-11, i.e. C shrinks against the ridge c*I.  Round-to-nearest is unbiased and its error averages
out over the n samples (relative 2^-12/sqrt(n) of the diagonal scale).  This script isolates the effect of the operand
rounding (products and sums exact, float64): weights from C(trunc(X)) and C(rn(X)) against weights from C(X), all
through the float64 oracle.

    python tools/next/emulate_tf32_rounding.py          # ~3 min on 8 cores
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
from oracle import restatement as R  # noqa: E402


def trunc_tf32(x):
    return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def rn_tf32(x):
    """cvt.rna.tf32.f32: round to nearest, ties away from zero, on the 13 dropped mantissa bits."""
    u = x.view(np.uint32).astype(np.uint64)
    u = (u + np.uint64(0x1000)) & np.uint64(0xFFFFE000)
    return u.astype(np.uint32).view(np.float32)


def cov64(views):
    X = np.hstack(views).astype(np.float64)
    n = X.shape[0]
    X -= X.mean(axis=0)
    return X.T @ X / (n - 1), n


def weights(views):
    C, n = cov64(views)
    w, sv = R.cov_rcca_fit(C, bench.DIMS, bench.K, bench.C_RIDGE, n)
    return w, sv


def err(w, w_ref):
    ws = R.align_signs(w, w_ref)
    per = np.concatenate([np.linalg.norm(a - b, axis=0) / np.linalg.norm(b, axis=0) for a, b in zip(ws, w_ref)])
    return per.max(), np.median(per)


if __name__ == "__main__":
    views = bench.make_views(1000)
    w_ref, sv = weights(views)
    print(f"reference spectrum: sigma_1 {sv[0]:.4f} sigma_64 {sv[63]:.4f}, min gap {np.min(-np.diff(sv[:64])):.2e}")
    for name, fn in (("truncated (what the tensor core does to raw fp32)", trunc_tf32), ("round-to-nearest", rn_tf32)):
        w, sv2 = weights([fn(np.ascontiguousarray(v)) for v in views])
        mx, med = err(w, w_ref)
        print(f"{name:52s}: max / median per-vector weight error {mx:.2e} / {med:.2e}; "
              f"canonical corr max rel err {np.max(np.abs(sv2[:64] - sv[:64]) / sv[:64]):.2e}")
