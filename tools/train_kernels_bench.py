import sys, torch
sys.path.insert(0, "/root/repo")
from vimoclip_b200 import ops
dev = torch.device("cuda:0")
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for rows, d in [(25600, 768), (25600, 640)]:
    z = torch.randn(rows, d, device=dev); dy = torch.randn(rows, d, device=dev); add = torch.randn(rows, d, device=dev)
    g = torch.ones(d, device=dev)
    print(rows, d, "layernorm_bwd(+add):", round(t(lambda: ops.layernorm_bwd(z, g, 1e-5, dy, add=add)), 4), "ms; no add:", round(t(lambda: ops.layernorm_bwd(z, g, 1e-5, dy)), 4), flush=True)
x = torch.randn(25600, 3072, device=dev); aux = torch.randn(25600, 3072, device=dev)
print("cast_colsum 25600x3072 qgelu:", round(t(lambda: ops.cast_colsum(x, aux)), 4), "ms; plain:", round(t(lambda: ops.cast_colsum(x)), 4))
