"""Launch the residual GEMM (out_proj shape) without and with the LayerNorm-fold producer outputs; for ncu captures."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vimoclip_b200 import ops
dev = torch.device("cuda:0")
F_, L, d = 1024, 197, 768
M = F_ * L
gen = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(M, d, device=dev, generator=gen).to(torch.bfloat16)
res = torch.randn(M, d, device=dev, generator=gen)
w = (torch.randn(d, d, device=dev, generator=gen) * d**-0.5).to(torch.bfloat16)
b = torch.randn(d, device=dev, generator=gen)
parts = ops.gemm_stats_parts(M, d)
raw16 = torch.empty(M, d, device=dev, dtype=torch.bfloat16)
stats = torch.zeros(parts, M, 2, device=dev)
for _ in range(int(os.environ.get("REPS", "2"))):
    ops.gemm(x, w, bias=b, resid=res, out=res)
    ops.gemm(x, w, bias=b, resid=res, out=res, emit_stats=(raw16, stats))
torch.cuda.synchronize()
