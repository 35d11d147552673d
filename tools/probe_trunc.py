"""Does tcgen05 kind::tf32 truncate or round its fp32 operands?  x = 1 + 0.75 ulp_tf32(1): truncation -> 1, RN -> 1 + 2^-10."""
import sys, torch
sys.path.insert(0, ".")
from cca_zoo_b200 import ops
n, d = 256, 128
for frac in [0.75, 0.25, 0.5]:
    x = torch.full((n, d), 1.0 + frac * 2.0 ** -10, device="cuda", dtype=torch.float32)
    mom = ops.moments([x], precision="tf32")
    C, _ = ops.covariance(mom, [d], n, center=False)
    m00 = float(C[0, 0]) * (n - 1) / n          # mean of products
    print(f"frac={frac}: mean product = {m00:.10f}  (trunc -> 1.0, RN -> {(1 + (2.0**-10 if frac >= 0.5 else 0))**2:.10f}, exact {(1 + frac * 2.0 ** -10)**2:.10f})")
