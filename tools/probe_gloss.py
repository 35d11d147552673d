"""Timing of GCCALoss forward+backward (CUDA events) at DGCCA-like sizes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cca_zoo_b200.deep import GCCALoss

for n, widths, dt in [(4096, [64, 64, 64], torch.float32), (4096, [64, 64, 64], torch.float64),
                      (1024, [16, 16, 16], torch.float32), (8192, [128, 128, 128, 128], torch.float32)]:
    zs = [torch.randn(n, w, device="cuda", dtype=dt).requires_grad_(True) for w in widths]
    fn = GCCALoss()
    for _ in range(3):
        fn(zs).backward()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        for z in zs:
            z.grad = None
        fn(zs).backward()
    b.record()
    torch.cuda.synchronize()
    print(f"GCCALoss n={n} widths={widths} {dt}: {a.elapsed_time(b) / 10:.3f} ms fwd+bwd")
