"""Would 3xTF32 keep its accuracy if the two cross terms (hi*lo, lo*hi) ran as bf16 x bf16 MMAs (kind::f16: twice the
TF32 rate, half the shared-memory bytes)?  Emulates one accumulator run (2048 samples) of X^T X entries with exact
fp64 accumulation, so only the operand formats differ:
    exact      : x*y
    tf32       : trunc(x)*trunc(y)
    3xtf32     : hi*hi + hi*lo' + lo'*hi           lo' = rna_tf32(x - hi)              (what K1 does today)
    tf32+bf16x : hi*hi + bf(hi)*bf(lo) + bf(lo)*bf(hi)        bf = round-to-nearest-even bf16

    python tools/next/emulate_x3_bf16_cross.py
"""
import numpy as np


def trunc_tf32(x):
    return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def rn_tf32(x):
    u = x.view(np.uint32).astype(np.uint64)
    return ((u + np.uint64(0x1000)) & np.uint64(0xFFFFE000)).astype(np.uint32).view(np.float32)


def rne_bf16(x):
    u = x.view(np.uint32).astype(np.uint64)
    u = (u + np.uint64(0x7FFF) + ((u >> np.uint64(16)) & np.uint64(1))) & np.uint64(0xFFFF0000)
    return u.astype(np.uint32).view(np.float32)


rng = np.random.default_rng(0)
n, d = 2048, 256
z = rng.standard_normal((n, 8)).astype(np.float32)
X = (z @ rng.standard_normal((8, d)).astype(np.float32) * 0.2 + rng.standard_normal((n, d)).astype(np.float32) * 22.6
     + 0.5).astype(np.float32)
hi = trunc_tf32(X)
lo = rn_tf32((X - hi).astype(np.float32))
X64, hi64, lo64 = X.astype(np.float64), hi.astype(np.float64), lo.astype(np.float64)
bhi, blo = rne_bf16(hi).astype(np.float64), rne_bf16(lo).astype(np.float64)
exact = X64.T @ X64
variants = {
    "tf32 (1 pass)": hi64.T @ hi64,
    "3xtf32": hi64.T @ hi64 + hi64.T @ lo64 + lo64.T @ hi64,
    "tf32 + bf16 cross terms": hi64.T @ hi64 + bhi.T @ blo + blo.T @ bhi,
}
scale = np.sqrt(np.outer(np.diag(exact), np.diag(exact)))
for name, M in variants.items():
    err = np.abs(M - exact) / scale
    print(f"{name:26s}: max {err.max():.2e}  rms {np.sqrt((err ** 2).mean()):.2e}  diagonal max {np.abs(np.diag(M - exact) / np.diag(exact)).max():.2e}")
