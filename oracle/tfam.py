"""fp32 CPU restatement of the TFAM fusion block (TEST INFRASTRUCTURE).

Follows ``TFAM/models/AMO_CLIP.py``: ``AttentionLayer`` (:6-51, post-LN: self-attn -> cross-attn ->
FFN with ReLU) and ``AMO_CLIP`` (:55-171: mask inversion :125-126, sinusoidal PE :88-97,129-134,
branch select :136-167, mean over ALL rows + classifier :170).  Module and parameter names are the
reference's, so ``state_dict`` keys match and reference checkpoints load.  The attention itself is
written out explicitly (no ``nn.MultiheadAttention.forward``) so the arithmetic is visible; the
``nn.MultiheadAttention`` modules are kept only as parameter containers.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def mha(attn: nn.MultiheadAttention, q_in, kv_in, key_padding_mask=None):
    """Multi-head attention, batch-first, eval mode.  key_padding_mask True = ignore."""
    d = attn.embed_dim
    h = attn.num_heads
    w, b = attn.in_proj_weight, attn.in_proj_bias
    q = F.linear(q_in, w[:d], b[:d])
    k = F.linear(kv_in, w[d : 2 * d], b[d : 2 * d])
    v = F.linear(kv_in, w[2 * d :], b[2 * d :])
    B, Tq, _ = q.shape
    Tk = k.shape[1]
    q = q.view(B, Tq, h, d // h).transpose(1, 2)
    k = k.view(B, Tk, h, d // h).transpose(1, 2)
    v = v.view(B, Tk, h, d // h).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(d // h)
    if key_padding_mask is not None:
        s = s.masked_fill(key_padding_mask[:, None, None, :], float("-inf"))
    o = torch.softmax(s, dim=-1) @ v
    o = o.transpose(1, 2).reshape(B, Tq, d)
    return F.linear(o, attn.out_proj.weight, attn.out_proj.bias)


class AttentionLayer(nn.Module):
    def __init__(self, d_model, num_heads, dim_feedforward, dropout=0.1, activation="relu"):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d_model, num_heads, dropout=dropout, batch_first=True)
        self.cross_attn = nn.MultiheadAttention(d_model, num_heads, dropout=dropout, batch_first=True)
        self.ffn = nn.Sequential(
            nn.Linear(d_model, dim_feedforward),
            nn.GELU() if activation == "gelu" else nn.ReLU(),
            nn.Dropout(dropout),
            nn.Linear(dim_feedforward, d_model),
            nn.Dropout(dropout),
        )
        self.norm_self = nn.LayerNorm(d_model)
        self.norm_cross = nn.LayerNorm(d_model)
        self.norm_ffn = nn.LayerNorm(d_model)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x, cross_src=None, src_key_padding_mask=None, cross_key_padding_mask=None):
        x = self.norm_self(x + mha(self.self_attn, x, x, src_key_padding_mask))  # :39-40
        if cross_src is not None:
            x = self.norm_cross(x + mha(self.cross_attn, x, cross_src, cross_key_padding_mask))  # :44-45
        h = self.ffn[3](F.relu(self.ffn[0](x)) if isinstance(self.ffn[1], nn.ReLU) else F.gelu(self.ffn[0](x)))
        return self.norm_ffn(x + h)  # :48-49


class TfamOracle(nn.Module):
    def __init__(
        self,
        d_model=512,
        nhead=8,
        num_layers=4,
        dim_feedforward=2048,
        num_classes=140,
        use_cross_attention=True,
        use_pe=False,
        use_only_rgb=False,
        use_only_flow=False,
        concat_dim=1,
        dropout=0.1,
        mlp_dropout=0.3,
        device="cpu",
    ):
        super().__init__()
        self.use_cross_attention = use_cross_attention
        self.use_pe = use_pe
        self.use_only_rgb = use_only_rgb
        self.use_only_flow = use_only_flow
        self.concat_dim = concat_dim
        self.d_model = d_model
        self.layers = nn.ModuleList([AttentionLayer(d_model, nhead, dim_feedforward, dropout=dropout) for _ in range(num_layers)])
        self.classifier = nn.Sequential(
            nn.LayerNorm(d_model), nn.Linear(d_model, d_model // 2), nn.GELU(), nn.Dropout(mlp_dropout), nn.Linear(d_model // 2, num_classes)
        )
        self.projection_layer = nn.Linear(2 * d_model, d_model)

    def positional_encoding(self, seq_len):  # :88-97
        pos = torch.arange(seq_len).unsqueeze(1)
        div = torch.exp(torch.arange(0, self.d_model, 2) * (-math.log(10000.0) / self.d_model))
        pe = torch.zeros(seq_len, self.d_model)
        pe[:, 0::2] = torch.sin(pos * div)
        pe[:, 1::2] = torch.cos(pos * div)
        return pe

    def forward(self, rgb_emb, motion_emb, mask_rgb=None, mask_flow=None):
        """Inference under no_grad; in ``.train()`` mode autograd stays on (the reference module has no decorator at
        all: TFAM/train_and_eval.py:80 backpropagates through it) -- the oracle of the training-step parity tests."""
        with torch.set_grad_enabled(self.training and torch.is_grad_enabled()):
            return self._forward(rgb_emb, motion_emb, mask_rgb, mask_flow)

    def _forward(self, rgb_emb, motion_emb, mask_rgb=None, mask_flow=None):
        attn_rgb = ~mask_rgb if mask_rgb is not None else None  # :125
        attn_flow = ~mask_flow if mask_flow is not None else None  # :126
        if self.use_pe:  # :129-134 (the reference adds in place; value semantics are the same)
            rgb_emb = rgb_emb + self.positional_encoding(rgb_emb.size(1)).unsqueeze(0)
            motion_emb = motion_emb + self.positional_encoding(motion_emb.size(1)).unsqueeze(0)
        if self.use_only_rgb:
            x = rgb_emb
            for layer in self.layers:
                x = layer(x, src_key_padding_mask=attn_rgb)
        elif self.use_only_flow:
            x = motion_emb
            for layer in self.layers:
                x = layer(x, src_key_padding_mask=attn_flow)
        elif self.use_cross_attention:
            x = rgb_emb
            for layer in self.layers:
                x = layer(x, cross_src=motion_emb, src_key_padding_mask=attn_rgb, cross_key_padding_mask=attn_flow)
        else:
            rgb_emb = rgb_emb[:, :-1, :]  # :153
            attn_rgb = attn_rgb[:, :-1]  # :154
            if self.concat_dim == 1:
                attn_mask = torch.cat([attn_rgb, attn_flow], dim=1)
                x = torch.cat([rgb_emb, motion_emb], dim=1)
            else:
                attn_mask = attn_flow
                x = self.projection_layer(torch.cat([rgb_emb, motion_emb], dim=-1))
            for layer in self.layers:
                x = layer(x, src_key_padding_mask=attn_mask)
        c = self.classifier
        return c[4](F.gelu(c[1](c[0](x.mean(dim=1)))))  # :170 mean over ALL rows incl. padding
