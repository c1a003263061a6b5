"""CPU emulation of the tower's rounding points, to decide the residual-stream dtype BEFORE spending GPU time.

Schemes (all: bf16 GEMM operands, fp32 accumulation, LayerNorm folded into the consuming GEMM as the library does):
  f32   residual stream fp32 (round 1), A operand = bf16(x), row statistics from the fp32 rows
  bf16  residual stream bf16: x = bf16(acc + bias + float(x)); row statistics from the rounded rows
Reports min cosine of the embeddings against the fp32 oracle and the TFAM logit error the embedding error causes.

    python tools/emulate_residual_precision.py [ViT-B/16] [clips]
"""
from __future__ import annotations

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import clip_shim, prologue, tfam as otfam, weights  # noqa: E402


def bf(t):
    return t.to(torch.bfloat16).float()


def tower(vit, x_img, resid_bf16: bool):
    d = vit.conv1.weight.shape[0]
    heads = d // 64
    p = vit.conv1.kernel_size[0]
    F = x_img.shape[0]
    g = 224 // p
    patches = x_img.reshape(F, 3, g, p, g, p).permute(0, 2, 4, 1, 3, 5).reshape(F, g * g, 3 * p * p)
    tok = bf(patches) @ bf(vit.conv1.weight.reshape(d, -1)).t() + vit.positional_embedding[1:]
    cls = (vit.class_embedding + vit.positional_embedding[0]).expand(F, 1, d)
    x = torch.cat([cls, tok], 1)
    x = torch.nn.functional.layer_norm(x, (d,), vit.ln_pre.weight, vit.ln_pre.bias, 1e-5)
    if resid_bf16:
        x = bf(x)

    def folded(x, ln, w, b):
        xr = bf(x)  # A operand (== x when the stream is bf16)
        src = xr if resid_bf16 else x
        mean = src.mean(-1, keepdim=True)
        var = (src * src).mean(-1, keepdim=True) - mean * mean
        rstd = torch.rsqrt(var + 1e-5)
        wf = bf(w * ln.weight[None, :])
        cs = wf.sum(1)
        bb = b + w @ ln.bias
        return rstd * (xr @ wf.t() - mean * cs) + bb

    for blk in vit.transformer.resblocks:
        qkv = bf(folded(x, blk.ln_1, blk.attn.in_proj_weight, blk.attn.in_proj_bias))
        q, k, v = [t.reshape(F, -1, heads, 64).transpose(1, 2) for t in qkv.split(d, -1)]
        s = (q @ k.transpose(-1, -2)) / 8.0
        pexp = bf(torch.exp(s - s.amax(-1, keepdim=True)))
        o = bf((pexp @ v) / pexp.sum(-1, keepdim=True))
        o = o.transpose(1, 2).reshape(F, -1, d)
        x = o @ bf(blk.attn.out_proj.weight).t() + blk.attn.out_proj.bias + x
        if resid_bf16:
            x = bf(x)
        h = folded(x, blk.ln_2, blk.mlp.c_fc.weight, blk.mlp.c_fc.bias)
        h = bf(h * torch.sigmoid(1.702 * h))
        x = h @ bf(blk.mlp.c_proj.weight).t() + blk.mlp.c_proj.bias + x
        if resid_bf16:
            x = bf(x)
    c = bf(torch.nn.functional.layer_norm(x[:, 0], (d,), vit.ln_post.weight, vit.ln_post.bias, 1e-5))
    return c @ bf(vit.proj)


@torch.no_grad()
def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "ViT-B/16"
    clips = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    T = 16
    torch.manual_seed(0)
    vit = clip_shim.build_visual(name, seed=0)
    gen = torch.Generator().manual_seed(1234)
    frames = torch.randint(0, 256, (clips * T, 3, 224, 224), dtype=torch.uint8, generator=gen)
    x = torch.from_numpy(prologue.normalise_u8(frames.numpy()))
    ref = vit(x)
    tf = otfam.TfamOracle().eval()
    weights.randomise_tfam_(tf, 0)
    mot = torch.randn(clips, T - 1, 512, generator=gen)
    out = {}
    for scheme in ("f32", "bf16"):
        e = tower(vit, x, scheme == "bf16")
        cos = torch.nn.functional.cosine_similarity(e.double(), ref.double(), dim=-1)
        rel = ((e - ref).norm(dim=-1) / ref.norm(dim=-1)).max().item()
        line = f"{name} {scheme:5s}: min cos {cos.min().item():.7f}  max rel err {rel:.3e}"
        if ref.shape[-1] == 512:
            lg_ref = tf(ref.view(clips, T, -1), mot)
            lg = tf(e.view(clips, T, -1), mot)
            line += f"  TFAM logit max-abs from RGB embedding error {(lg - lg_ref).abs().max().item():.3e}"
        print(line, flush=True)
        out[scheme] = e


if __name__ == "__main__":
    main()
