"""Per-tile timeline of the tcgen05 GEMM's CTA 0 (clock64 stamps through VMC_OPT_DEBUG_PTR): when the MMA issuer starts and
finishes issuing a tile, when epilogue warp 0 gets the accumulator and when it and warp 7 have drained it.  For tuning."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vimoclip_b200 as vmc  # noqa: E402
from vimoclip_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
M, d = int(os.environ.get("GS_FRAMES", "2048")) * 197, 768
gen = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(M, d, device=dev, generator=gen).to(torch.bfloat16)
h = torch.randn(M, 4 * d, device=dev, generator=gen).to(torch.bfloat16)
xs = torch.randn(M, d, device=dev, generator=gen).to(torch.bfloat16)
parts = ops.gemm_stats_parts(M, d)
stats = torch.zeros(parts, M, 2, device=dev)
xf = x.float()
stats_in = torch.zeros(1, M, 2, device=dev)
stats_in[0, :, 0] = xf.sum(1)
stats_in[0, :, 1] = (xf * xf).sum(1)
del xf
ops.set_option(vmc._lib.OPT_GEMM_IMPL, int(os.environ.get("GS_IMPL", "0")))
names = ["mma:start", "mma:issued", "ep0:wait", "ep0:acc_ready", "ep0:done", "ep7:done"]


def case(name, a, N, K, **kw):
    w = (torch.randn(N, K, device=dev, generator=gen) * K**-0.5).to(torch.bfloat16)
    b = None if kw.pop("nobias", False) else torch.randn(N, device=dev, generator=gen)
    if kw.pop("fold", False):
        kw["fold"] = (stats_in, torch.randn(N, device=dev, generator=gen), 1e-5)
    out = kw.pop("out", None)
    if out is None:
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(20):  # reach the power-capped steady state
        ops.gemm(a, w, bias=b, out=out, **kw)
    dbg = torch.zeros(32 * 8, dtype=torch.int64, device=dev)
    ops.set_option(7, dbg.data_ptr())
    ops.gemm(a, w, bias=b, out=out, **kw)
    torch.cuda.synchronize()
    ops.set_option(7, 0)
    t = dbg.cpu().view(32, 8)
    t0 = int(t[8, 0])
    print(f"== {name}: N={N} K={K}")
    for i in range(8, 16):
        print(f"  tile {i}: " + "  ".join(f"{n}@{int(t[i, s]) - t0}" for s, n in enumerate(names)))
    per = (int(t[24, 0]) - int(t[8, 0])) / 16.0
    ep = sum(int(t[i, 4]) - int(t[i, 3]) for i in range(8, 24)) / 16.0
    ep7 = sum(int(t[i, 5]) - int(t[i, 3]) for i in range(8, 24)) / 16.0
    iss = sum(int(t[i, 1]) - int(t[i, 0]) for i in range(8, 24)) / 16.0
    wait = sum(int(t[i, 3]) - int(t[i, 2]) for i in range(8, 24)) / 16.0
    starved = sum(int(t[i, 6]) for i in range(8, 24)) / 16.0
    print(f"  per tile: period {per:.0f} cycles, MMA issue span {iss:.0f} (of which waiting for operands {starved:.0f}), epilogue warp 0 {ep:.0f} (warp 7 {ep7:.0f}), epilogue idle before the accumulator {wait:.0f}", flush=True)


n_tok = 196
pe_rows = (M // 197) * n_tok
pos = torch.randn(197, d, device=dev, generator=gen)
pe_out = torch.zeros((M // 197) * 197, d, device=dev)
if os.environ.get("GS_PATCH", "1") == "1":
    case("patch embed (generic: row remap + pos-emb, fp32 out)", x[:pe_rows], d, d, resid=pos, out=pe_out, row_group=n_tok, nobias=True)
case("qkv LN-fold", x, 3 * d, d, fold=True)
case("qkv plain", x, 3 * d, d)
case("c_fc LN-fold+QuickGELU", x, 4 * d, d, fold=True, act=ops.ACT_QUICKGELU)
case("c_fc plain (no act)", x, 4 * d, d)
case("out_proj shape, plain bias epilogue", x, d, d)
case("out_proj bf16-resid+stats", x, d, d, resid=xs, out=xs, emit_stats=(None, stats))
case("c_proj bf16-resid+stats", h, d, 4 * d, resid=xs, out=xs, emit_stats=(None, stats))
