import sys, torch
sys.path.insert(0, ".")
from cca_zoo_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
n = 1024
X = torch.randn(2, 2 * n, n, generator=g, device="cuda")
A0 = X.transpose(1, 2) @ X / (2 * n) + 0.1 * torch.eye(n, device="cuda")
for _ in range(2):
    A = A0.clone(); ops.potrf_inv_(A)
A = torch.randn(1024, 1024, generator=g, device="cuda"); Z = torch.randn(1024, 96, generator=g, device="cuda")
for _ in range(2):
    ops.gemm_tc(A, A); ops.gemm_tc(A, Z); ops.gemm_tc(Z, Z, transa=True); ops.gemm_tc(Z, A[:96, :96].contiguous())
H = Z.T @ Z
for _ in range(2):
    ops.syevj_small(H)
torch.cuda.synchronize()
