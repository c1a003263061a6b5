"""CPU emulation: logit error of the TFAM block when every Linear's operands are rounded to bf16 / fp16 (fp32 accumulation),
against the fp32 oracle -- the evidence behind the fused kernel's fp16 operands (csrc/tfam_fused.cu).

    python tools/emulate_tfam_precision.py     ->  bf16 ~1.2e-2 (over the 1e-2 bar), fp16 ~1.7e-3 (max 2.4e-3)
"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle.tfam as T  # noqa: E402
from oracle import tfam as otfam, weights  # noqa: E402

orig_linear = F.linear
quant = {"bf16": lambda t: t.to(torch.bfloat16).float(), "fp16": lambda t: t.to(torch.float16).float()}
MODES = [dict(), dict(use_only_rgb=True), dict(use_only_flow=True), dict(use_cross_attention=False, concat_dim=1),
         dict(use_cross_attention=False, concat_dim=-1), dict(use_pe=True)]


def main():
    res = {k: 0.0 for k in quant}
    with torch.no_grad():
        for seed in range(6):
            for kw in MODES:
                m = otfam.TfamOracle(**kw).eval()
                weights.randomise_tfam_(m, seed)
                g = torch.Generator().manual_seed(100 + seed)
                B = 8
                rgb, mot = torch.randn(B, 16, 512, generator=g), torch.randn(B, 15, 512, generator=g)
                lr = torch.randint(8, 17, (B,), generator=g)
                lm = torch.clamp(lr - 1, max=15)
                mr, mm = torch.arange(16)[None] < lr[:, None], torch.arange(15)[None] < lm[:, None]
                T.F.linear = orig_linear
                ref = m(rgb.clone(), mot.clone(), mr, mm)
                for name, q in quant.items():
                    T.F.linear = lambda x, w, b=None, q=q: orig_linear(q(x), q(w), b)
                    res[name] = max(res[name], (m(rgb.clone(), mot.clone(), mr, mm) - ref).abs().max().item())
                T.F.linear = orig_linear
    print({k: f"{v:.3e}" for k, v in res.items()})


if __name__ == "__main__":
    main()
