"""Per-shape timing of the hot kernels with CUDA events (not part of bench.py; for tuning)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vimoclip_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
F_ = int(os.environ.get("KB_FRAMES", "1024"))
which = os.environ.get("KB_WHICH", "gemm,attn,ln").split(",")
iters = int(os.environ.get("KB_ITERS", "10"))


def timeit(fn, iters=iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


gen = torch.Generator(device="cuda").manual_seed(0)
import vimoclip_b200 as vmc  # noqa: E402
ops.set_option(vmc._lib.OPT_GEMM_IMPL, int(os.environ.get("KB_GEMM_IMPL", "0")))
ops.set_option(vmc._lib.OPT_ATTN_PREFETCH, int(os.environ.get("KB_ATTN_PREFETCH", "0")))
if "gemm" in which:
    for L, d in [(197, 768), (50, 768)]:
        M = F_ * L
        x = torch.randn(M, d, device=dev, generator=gen).to(torch.bfloat16)
        h = torch.randn(M, 4 * d, device=dev, generator=gen).to(torch.bfloat16)
        res = torch.randn(M, d, device=dev, generator=gen)
        for name, a, N, K, kw in [
            ("qkv", x, 3 * d, d, dict(out_dtype=torch.bfloat16)),
            ("out+res", x, d, d, dict(resid=res, out=res)),
            ("fc1+qgelu", x, 4 * d, d, dict(act=ops.ACT_QUICKGELU, out_dtype=torch.bfloat16)),
            ("fc1 plain", x, 4 * d, d, dict(out_dtype=torch.bfloat16)),
            ("fc2+res", h, d, 4 * d, dict(resid=res, out=res)),
            ("fc2 plain bf16", h, d, 4 * d, dict(out_dtype=torch.bfloat16)),
        ]:
            w = (torch.randn(N, K, device=dev, generator=gen) * K**-0.5).to(torch.bfloat16)
            b = torch.randn(N, device=dev, generator=gen)
            out = kw.pop("out", None)
            if out is None:
                out = torch.empty(M, N, device=dev, dtype=kw.pop("out_dtype"))
            ms = timeit(lambda: ops.gemm(a, w, bias=b, out=out, **kw))
            print(f"gemm {name:16s} M={M} N={N} K={K}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s", flush=True)
        # LayerNorm folding: producers emit bf16 rows + row statistics, consumers apply mean / rstd in the epilogue
        parts = ops.gemm_stats_parts(M, d)
        raw16 = torch.empty(M, d, device=dev, dtype=torch.bfloat16)
        stats = torch.zeros(parts, M, 2, device=dev)
        for name, a, N, K in [("out+res+stats", x, d, d), ("fc2+res+stats", h, d, 4 * d)]:
            w = (torch.randn(N, K, device=dev, generator=gen) * K**-0.5).to(torch.bfloat16)
            b = torch.randn(N, device=dev, generator=gen)
            ms = timeit(lambda: ops.gemm(a, w, bias=b, resid=res, out=res, emit_stats=(raw16, stats)))
            print(f"gemm {name:16s} M={M} N={N} K={K}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s", flush=True)
        stats.normal_().abs_()
        stats[..., 1] += 100.0
        for name, N, act in [("qkv fold", 3 * d, ops.ACT_NONE), ("fc1+qgelu fold", 4 * d, ops.ACT_QUICKGELU)]:
            w = (torch.randn(N, d, device=dev, generator=gen) * d**-0.5).to(torch.bfloat16)
            b = torch.randn(N, device=dev, generator=gen)
            cs = w.float().sum(1)
            out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            ms = timeit(lambda: ops.gemm(x, w, bias=b, act=act, out=out, fold=(stats, cs, 1e-5)))
            print(f"gemm {name:16s} M={M} N={N} K={d}: {ms:.3f} ms  {2.0 * M * N * d / ms / 1e9:.0f} TFLOP/s", flush=True)
if "attn" in which:
    for L, heads in [(197, 12), (50, 12), (257, 16)][: int(os.environ.get("KB_ATTN_SHAPES", "3"))]:
        d = heads * 64
        Fa = F_ if L != 257 else F_ // 2
        qkv = torch.randn(Fa * L, 3 * d, device=dev, generator=gen).to(torch.bfloat16)
        for impl in [int(x) for x in os.environ.get("KB_ATTN_IMPLS", "5,2").split(",")]:
            ms = timeit(lambda: ops.attention_vit(qkv, Fa, L, heads, impl=impl))
            fl = 4.0 * Fa * heads * L * L * 64
            print(f"attn impl={impl} L={L} heads={heads} F={Fa}: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s  "
                  f"{ms * 1e3 * 148 / (Fa * heads):.1f} us per (frame,head) per SM", flush=True)
if "ln" in which:
    x = torch.randn(F_ * 197, 768, device=dev, generator=gen)
    g_ = torch.ones(768, device=dev)
    ms = timeit(lambda: ops.layernorm(x, g_, g_))
    print(f"layernorm rows={F_ * 197} d=768: {ms:.3f} ms  {x.numel() * 6 / ms / 1e6:.0f} GB/s")
    u8 = torch.randint(0, 256, (F_, 3, 224, 224), dtype=torch.uint8, device=dev, generator=gen)
    for impl in (0, 1, 2):
        ops.set_option(vmc._lib.OPT_PROLOGUE_IMPL, impl)
        for p in (16, 32, 14):
            ms = timeit(lambda: ops.prologue(u8, wrap=True, dst="patch", patch=p))
            print(f"prologue impl={impl} p={p} F={F_}: {ms:.3f} ms  {u8.numel() * 3 / ms / 1e6:.0f} GB/s")
    bgr = torch.randint(0, 256, (F_ // 16, 17, 224, 224, 3), dtype=torch.uint8, device=dev, generator=gen)
    for impl in (0, 1, 2):
        ops.set_option(vmc._lib.OPT_PROLOGUE_IMPL, impl)
        ms = timeit(lambda: ops.frame_diff(bgr, dst="patch", patch=32, want_diff=False))
        nfr = (F_ // 16) * 16
        # algorithmic bytes: every BGR frame once (17 per 16 outputs, 3 B/px) + 3 bf16 channels written (6 B/px)
        print(f"frame_diff+prologue impl={impl} p=32 frames={nfr}: {ms:.3f} ms  "
              f"{nfr * 50176 * (3 * 17 / 16 + 6) / ms / 1e6:.0f} GB/s")
