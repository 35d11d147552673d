#!/bin/bash
echo "== probe =="; timeout 120 python tools/probe_tf32.py 2>&1 | tail -6
timeout 300 python - <<'PY'
import sys, torch
sys.path.insert(0, ".")
from cca_zoo_b200 import ops, _lib
lib = _lib.load()
def timeit(fn, warm=3, it=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ks=[]
    for _ in range(it):
        fn(); ks.append(lib.ccab_profile_moments_last_ms())
    return sum(ks)/len(ks), min(ks)
torch.manual_seed(0)
n, d = 100000, 1024
X1 = torch.randn(n, d, device="cuda"); X2 = torch.randn(n, d, device="cuda")
F = n * 2 * d * (2 * d + 1)
lib.ccab_profile_moments(1)
for variant, kc in [(0, 32), (0, 64), (0, 16), (1, 32)]:
    ops.debug_set("tc_variant", variant); ops.debug_set("tc_kc", kc)
    for prec in (["tf32", "tf32x3"] if kc == 32 else ["tf32"]):
        k, kmin = timeit(lambda: ops.moments([X1, X2], precision=prec))
        print(f"variant={variant} kc={kc} {prec}: kernel avg {k:.3f} ms min {kmin:.3f} -> {F/k/1e9:.1f} TFLOP/s algorithmic", flush=True)
ops.debug_set("tc_variant", 0); ops.debug_set("tc_kc", 32); ops.debug_set("tc_dry_run", 1)
k, kmin = timeit(lambda: ops.moments([X1, X2], precision="tf32")); print(f"dry kc=32: {k:.3f}")
PY
echo "== tests =="; timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "moments" 2>&1 | tail -2
