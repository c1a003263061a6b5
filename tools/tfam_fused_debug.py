"""Fused TFAM kernel: error against the fp32 oracle / the batched path for a list of shapes, then latency and throughput.

    python tools/tfam_fused_debug.py            (GPU box)
Prints one line per case and never stops at the first failure (GPU minutes are scarce)."""
from __future__ import annotations

import os
import sys
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vimoclip_b200 as vmc  # noqa: E402
from oracle import tfam as otfam, weights  # noqa: E402
from vimoclip_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
MODES = {"cross": dict(), "rgb_only": dict(use_only_rgb=True), "flow_only": dict(use_only_flow=True),
         "concat_t": dict(use_cross_attention=False, concat_dim=1), "concat_e": dict(use_cross_attention=False, concat_dim=-1)}


def pair(mode, seed=0):
    o = otfam.TfamOracle(**MODES[mode]).eval()
    weights.randomise_tfam_(o, seed)
    m = vmc.AMO_CLIP(device=dev, **MODES[mode])
    m.load_state_dict(o.state_dict(), strict=True)
    return o, m.to(dev).eval()


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    for mode in MODES:
        o, m = pair(mode)
        for B, T, Tm in [(2, 16, 15), (1, 16, 16), (33, 16, 15), (256, 16, 15), (5, 32, 31), (7, 9, 4)]:
            if mode.startswith("concat") and T != Tm + 1:
                continue
            try:
                gen = torch.Generator().manual_seed(B + T)
                rgb, mot = torch.randn(B, T, 512, generator=gen), torch.randn(B, Tm, 512, generator=gen)
                lr, lm = torch.randint(1, T + 1, (B,), generator=gen), torch.randint(1, Tm + 1, (B,), generator=gen)
                if mode.startswith("concat"):
                    lr = torch.clamp(lr, min=2)
                mr, mm = torch.arange(T)[None] < lr[:, None], torch.arange(Tm)[None] < lm[:, None]
                ref = o(rgb.clone(), mot.clone(), mr, mm)
                args = [t.to(dev) for t in (rgb, mot, mr, mm)]
                m.fused = True
                ops.reset_launch_count()
                f = m(args[0].clone(), args[1].clone(), args[2], args[3])
                nl = ops.launch_count()
                m.fused = False
                b = m(args[0].clone(), args[1].clone(), args[2], args[3])
                torch.cuda.synchronize()
                ef, eb = (f.cpu() - ref).abs().max().item(), (b.cpu() - ref).abs().max().item()
                bad = (f.cpu() - ref).abs().amax(dim=1)
                print(f"{mode:9s} B={B:3d} T={T:2d} Tm={Tm:2d}: fused err {ef:.3e} ({nl} launches)  batched err {eb:.3e}  "
                      f"worst clips {torch.topk(bad, min(3, B)).indices.tolist()} nan={bool(torch.isnan(f).any())}", flush=True)
            except Exception:
                print(f"{mode} B={B} T={T} Tm={Tm}: EXCEPTION\n{traceback.format_exc()}", flush=True)
    # ---- latency / throughput ----
    o, m = pair("cross")
    for B in (2, 8, 32, 256):
        gen = torch.Generator().manual_seed(B)
        rgb, mot = torch.randn(B, 16, 512, generator=gen).to(dev), torch.randn(B, 15, 512, generator=gen).to(dev)
        mr = (torch.arange(16)[None] < torch.randint(8, 17, (B, 1), generator=gen)).to(dev)
        mm = (torch.arange(15)[None] < torch.randint(8, 16, (B, 1), generator=gen)).to(dev)
        res = {}
        for fused in (True, False):
            m.fused = fused
            res["fused" if fused else "batched"] = timeit(lambda: m(rgb, mot, mr, mm))
            try:
                gm = vmc.graphed(m, rgb, mot, mr, mm)
                res[("fused" if fused else "batched") + "_graph"] = timeit(lambda: gm(rgb, mot, mr, mm))
            except Exception as e:  # noqa: BLE001
                res[("fused" if fused else "batched") + "_graph"] = f"failed: {e}"
        print(f"TFAM cross B={B:3d} T=16/15 ms per forward: " + "  ".join(f"{k} {v if isinstance(v, str) else format(v, '.4f')}" for k, v in res.items()), flush=True)


if __name__ == "__main__":
    main()
