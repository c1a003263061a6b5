"""Diagnostics for a first run on the B200: per-kernel error patterns (not a test; prints only)."""
import os
import sys
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vimoclip_b200 as vmc  # noqa: E402
from vimoclip_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
print("device", torch.cuda.get_device_name(0), ops.device_info())


def run(name, fn):
    try:
        fn()
        torch.cuda.synchronize()
    except Exception:
        print(f"[{name}] EXCEPTION")
        traceback.print_exc()


def gemm_case(M, N, K, odt=torch.float32):
    gen = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn(M, K, device=dev, generator=gen).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev, generator=gen) * K**-0.5).to(torch.bfloat16)
    got = ops.gemm(a, w, out_dtype=odt).float()
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t()
    err = (got - ref).abs()
    print(f"[gemm {M}x{N}x{K} {odt}] max err {err.max().item():.4e} ref max {ref.abs().max().item():.3f} "
          f"bad rows {(err.max(1).values > 0.05).sum().item()} bad cols {(err.max(0).values > 0.05).sum().item()} "
          f"nan {torch.isnan(got).sum().item()}")
    if err.max().item() > 0.05:
        r = err.max(1).values
        c = err.max(0).values
        print("   first bad rows", (r > 0.05).nonzero().flatten()[:16].tolist())
        print("   first bad cols", (c > 0.05).nonzero().flatten()[:16].tolist())
        print("   got[0,:8]", got[0, :8].tolist())
        print("   ref[0,:8]", ref[0, :8].tolist())


def attn_case(L, heads, F_):
    gen = torch.Generator(device="cuda").manual_seed(L)
    d = heads * 64
    qkv = (torch.randn(F_ * L, 3 * d, device=dev, generator=gen) * 1.5).to(torch.bfloat16)
    got = ops.attention_vit(qkv, F_, L, heads).float().view(F_, L, heads, 64)
    torch.cuda.synchronize()
    q, k, v = qkv.float().view(F_, L, 3, heads, 64).unbind(2)
    s = torch.einsum("flhd,fmhd->fhlm", q, k) / 8.0
    ref = torch.einsum("fhlm,fmhd->flhd", torch.softmax(s, -1), v)
    err = (got - ref).abs()
    print(f"[attn L={L} h={heads} F={F_}] max err {err.max().item():.4e} nan {torch.isnan(got).sum().item()} "
          f"bad rows {(err.amax((0, 2, 3)) > 0.05).nonzero().flatten()[:16].tolist()}")


for shp in [(128, 128, 64), (128, 128, 128), (128, 128, 512), (256, 256, 64), (300, 512, 768), (40000, 768, 768)]:
    run("gemm", lambda shp=shp: gemm_case(*shp))
run("gemm bf16", lambda: gemm_case(1000, 2304, 768, torch.bfloat16))
for c in [(16, 1, 1), (50, 12, 2), (128, 2, 2), (197, 12, 2), (257, 16, 1)]:
    run("attn", lambda c=c: attn_case(*c))
print("launches", ops.launch_count())
