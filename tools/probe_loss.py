import sys, time, torch
sys.path.insert(0, ".")
from cca_zoo_b200.deep import CCALoss
torch.manual_seed(0)
for (B, k) in [(4096, 64), (4096, 512)]:
    zl = torch.randn(B, 16, device="cuda")
    z = [(zl @ torch.randn(16, k, device="cuda") + torch.randn(B, k, device="cuda")).requires_grad_(True) for _ in range(2)]
    fn = CCALoss(eps=1e-5)
    def step():
        for t in z: t.grad = None
        l = fn(z); l.backward(); return l
    for _ in range(3): step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): l = step()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20 * 1e3
    print(f"CCALoss fwd+bwd B={B} k={k}: {dt:.3f} ms  loss={l.item():.4f}")
