"""torch.matmul (cuBLAS) on the tower's GEMM shapes, twice each: for an ncu pass that shows which kernels, grids, cluster
shapes and resources cuBLAS picks (comparison point of profiles/r02_gemm_sustained_per_shape_clocks.txt)."""
import torch
dev = torch.device("cuda:0")
M, d = 2048 * 197, 768
x = torch.randn(M, d, device=dev).to(torch.bfloat16)
h = torch.randn(M, 4 * d, device=dev).to(torch.bfloat16)
for a, N, K in [(x, 3 * d, d), (x, 4 * d, d), (x, d, d), (h, d, 4 * d)]:
    w = torch.randn(N, K, device=dev).to(torch.bfloat16)
    o = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(2):
        torch.matmul(a, w.t(), out=o)
torch.cuda.synchronize()
print("ok")
