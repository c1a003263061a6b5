"""Sustained (power-capped) throughput of our tcgen05 GEMM vs torch.matmul (cuBLAS) on the four GEMM shapes of a
ViT-B/16 block at 2048 frames in flight.  Each case runs back to back for ~1.5 s; CUDA events around the whole run.
cuBLAS gets the plain product only (no bias / activation / residual): a lower bound on the work our epilogues do."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vimoclip_b200 import ops
dev = torch.device("cuda:0")
M, d = 2048 * 197, 768
gen = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(M, d, device=dev, generator=gen).to(torch.bfloat16)
h = torch.randn(M, 4 * d, device=dev, generator=gen).to(torch.bfloat16)
res = torch.randn(M, d, device=dev, generator=gen)


def sustained(fn, seconds=1.5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n, t0 = 0, time.perf_counter()
    e0.record()
    while time.perf_counter() - t0 < seconds:
        for _ in range(10):
            fn()
        n += 10
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for name, a, N, K, kw in [("qkv", x, 3 * d, d, dict(out_dtype=torch.bfloat16)),
                          ("out+res", x, d, d, dict(resid=res, out=res)),
                          ("fc1+qgelu", x, 4 * d, d, dict(act=ops.ACT_QUICKGELU, out_dtype=torch.bfloat16)),
                          ("fc2+res", h, d, 4 * d, dict(resid=res, out=res))]:
    w = (torch.randn(N, K, device=dev, generator=gen) * K**-0.5).to(torch.bfloat16)
    b = torch.randn(N, device=dev, generator=gen)
    out = kw.pop("out", None)
    if out is None:
        out = torch.empty(M, N, device=dev, dtype=kw.pop("out_dtype"))
    ms_ours = sustained(lambda: ops.gemm(a, w, bias=b, out=out, **kw))
    wt = w.t().contiguous()
    o2 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    ms_cublas = sustained(lambda: torch.matmul(a, wt, out=o2))
    ms_cublas_nt = sustained(lambda: torch.matmul(a, w.t(), out=o2))
    fl = 2.0 * M * N * K / 1e9
    print(f"{name:10s} M={M} N={N} K={K}: ours (fused epilogue) {fl / ms_ours:.0f} TFLOP/s | cuBLAS plain bf16 {fl / ms_cublas:.0f} (NN) "
          f"{fl / ms_cublas_nt:.0f} (NT) TFLOP/s", flush=True)
