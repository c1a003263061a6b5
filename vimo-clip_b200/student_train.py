"""Student training step: forward + backward of ``FlowStudentModel`` / ``FrameDiffStudentModel`` in ``.train()`` mode.

Reference: ``train.py:86-107`` -- ``embeddings, embeddings_distill, logits = model(flow_videos)``,
``loss = distillation_loss(embeddings_distill, teacher[:, :-1]) + classification_loss(logits, labels)``,
``loss.backward()``, ``Adam.step()`` on ALL parameters including the ViT (``train.py:66``).  The drop-in keeps that call
sequence: in training mode the forward returns the three outputs attached to autograd through
:class:`StudentTrainFunction`, whose backward returns the gradient of every parameter (tower + heads).

Forward = the inference kernels, orchestrated here layer by layer so the activations the backward needs stay in HBM
(x, ln_1(x), qkv, attention output, x after out_proj, ln_2, the c_fc pre-activation, QuickGELU(.)).  Backward:
  * every Linear / the patch-embedding conv: ``dX = dY W`` and ``dW = dY^T X`` as bf16 tcgen05 GEMMs (fp32 accumulation);
    the row-major activations / weights are read as MN-major UMMA operands (``vmc_gemm_bf16_ex``: no transposing pass);
    bias gradients are deterministic column sums;
  * LayerNorm backward from the saved LayerNorm inputs; QuickGELU backward element-wise;
  * attention backward (``vmc_attention_masked_bwd``): one CTA per (frame, head) with the probabilities recomputed in shared
    memory for ViT-B/32 (50 tokens, the reference's training default, ``train.py:63``), tiled flash-style kernels for
    longer sequences (ViT-B/16: 197 tokens);
  * heads (ResidualMLP, temporal mean, classification head) in split-bf16 GEMMs like the TFAM step.
Operands are bf16 (activations and gradients rounded once per GEMM), accumulators fp32: gradients agree with fp32 autograd of
the reference to bf16 accuracy (a few 1e-3 .. 1e-2 relative per tensor; tests/test_gpu_parity.py).
"""
from __future__ import annotations

import torch

from . import ops
from ._params import named_tensors
from .tfam_train import _activate, _lin_bwd, _lin_fwd

_PER_BLOCK = ("ln_1.weight", "ln_1.bias", "attn.in_proj_weight", "attn.in_proj_bias", "attn.out_proj.weight", "attn.out_proj.bias",
              "ln_2.weight", "ln_2.bias", "mlp.c_fc.weight", "mlp.c_fc.bias", "mlp.c_proj.weight", "mlp.c_proj.bias")
_TOWER_HEAD = ("conv1.weight", "class_embedding", "positional_embedding", "ln_pre.weight", "ln_pre.bias")
_TOWER_TAIL = ("ln_post.weight", "ln_post.bias", "proj")
_HEADS = ("residual_mlp.fc1.weight", "residual_mlp.fc1.bias", "residual_mlp.fc2.weight", "residual_mlp.fc2.bias",
          "classification_head.0.weight", "classification_head.0.bias", "classification_head.2.weight", "classification_head.2.bias")


def trainable_parameters(model):
    named = dict(named_tensors(model))  # also on DataParallel replicas (train.py:64), whose _parameters are empty
    n_layers = model.visual_encoder.layers
    names = [f"visual_encoder.{n}" for n in _TOWER_HEAD]
    names += [f"visual_encoder.transformer.resblocks.{i}.{n}" for i in range(n_layers) for n in _PER_BLOCK]
    names += [f"visual_encoder.{n}" for n in _TOWER_TAIL] + list(_HEADS)
    return names, [named[n] for n in names]


def _bf(t):
    return t.detach().to(torch.bfloat16).contiguous()


def _gemm16(a16, w16, **kw):
    return ops.gemm(a16, w16, **kw)


def _lin_bwd16(dy32, x16, w32, need_dx=True, qgelu_pre=None):
    """bf16 Linear backward: y = x W^T + b with x16 [M,K] bf16 (saved), W [N,K] fp32 parameter, dy fp32 [M,N].
    Returns (dx fp32 | None, dW fp32 [N,K], db fp32 [N]).  With ``qgelu_pre`` (the layer's pre-activation, fp32 [M,N]) ``dy32`` is
    the gradient AFTER the QuickGELU and the activation backward is applied in the same pass that casts it and sums its columns."""
    M, N = dy32.shape
    if N % 8 == 0 and dy32.is_contiguous():
        dy16, db = ops.cast_colsum(dy32, qgelu_pre)  # one pass: (activation backward,) bf16 operand, bias gradient
        dx = _gemm16(dy16, _bf(w32), w_t=True, out_dtype=torch.float32) if need_dx else None
        return dx, _gemm16(dy16, x16, a_t=True, w_t=True, out_dtype=torch.float32), db
    if qgelu_pre is not None:
        dy32 = ops.eltwise(ops.ELT_QGELU_BWD, dy32, qgelu_pre)
    dy16 = ops.cast_bf16(dy32)[:, :N]  # (row stride padded to a multiple of 8 when N is not one)
    dx = None
    if need_dx:  # dx[M,K] = dy[M,N] W[N,K]: W [N,K] row-major IS the transposed "weight" operand (MN-major B)
        dx = _gemm16(dy16, _bf(w32), w_t=True, out_dtype=torch.float32)
    # dW[N,K] = dy^T x: both operands are the row-major activations, read as MN-major UMMA operands (no transposing pass)
    dw = _gemm16(dy16, x16, a_t=True, w_t=True, out_dtype=torch.float32)
    return dx, dw, ops.colsum(dy32)


class StudentTrainFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, patches, *params):
        p, d, heads, n_layers, F_, B, T = cfg["patch"], cfg["d"], cfg["heads"], cfg["layers"], cfg["F"], cfg["B"], cfg["T"]
        g = cfg["res"] // p
        n, L = g * g, g * g + 1
        M = F_ * L
        kp = 3 * p * p
        conv_w, cls_emb, pos, gpre, bpre = params[:5]
        nb = len(_PER_BLOCK)
        blocks = [params[5 + i * nb:5 + (i + 1) * nb] for i in range(n_layers)]
        gpost, bpost, proj = params[5 + n_layers * nb:5 + n_layers * nb + 3]
        fc1w, fc1b, fc2w, fc2b, c1w, c1b, c2w, c2b = params[5 + n_layers * nb + 3:]
        dev = patches.device
        # patch embedding as a GEMM (+ positional embedding of the token rows), CLS rows = class_embedding + pos[0]
        ldp = patches.shape[1]
        wp = torch.zeros((d, ldp), dtype=torch.bfloat16, device=dev)
        wp[:, :kp] = conv_w.detach().reshape(d, kp).to(torch.bfloat16)
        z0 = torch.empty((M, d), dtype=torch.float32, device=dev)
        ops.gemm(patches, wp, resid=pos.detach().float().contiguous(), out=z0, k=kp, row_group=n)
        z0.view(F_, L, d)[:, 0, :] = (cls_emb.detach() + pos.detach()[0]).float()
        x, _ = ops.layernorm(z0, gpre.detach().float(), bpre.detach().float(), want32=True, want16=False)
        saved = []
        for (g1, b1_, wqkv, bqkv, wo, bo, g2, b2_, w1, bb1, w2, bb2) in blocks:
            _, xn = ops.layernorm(x, g1.detach().float(), b1_.detach().float())
            qkv = _gemm16(xn, _bf(wqkv), bias=bqkv.detach().float(), out_dtype=torch.bfloat16)
            a = ops.attention_vit(qkv, F_, L, heads)
            x_mid = _gemm16(a, _bf(wo), bias=bo.detach().float(), resid=x, out_dtype=torch.float32)
            _, xn2 = ops.layernorm(x_mid, g2.detach().float(), b2_.detach().float())
            h_pre = _gemm16(xn2, _bf(w1), bias=bb1.detach().float(), out_dtype=torch.float32)
            h16 = ops.qgelu_cast(h_pre)
            x_next = _gemm16(h16, _bf(w2), bias=bb2.detach().float(), resid=x_mid, out_dtype=torch.float32)
            saved.append(dict(x=x, xn=xn, qkv=qkv, a=a, x_mid=x_mid, xn2=xn2, h_pre=h_pre, h16=h16))
            x = x_next
        z_cls = x.view(F_, L, d)[:, 0, :].contiguous()  # ln_post input: the CLS rows
        _, cls16 = ops.layernorm(z_cls, gpost.detach().float(), bpost.detach().float())
        emb = _gemm16(cls16, _bf(proj.detach().t()), out_dtype=torch.float32)  # [F, D]
        D = emb.shape[1]
        # heads (models/student_model.py:90-96): distill = emb + alpha * fc2(GELU(fc1(emb))); logits = head(mean_T(emb))
        r_pre = _lin_fwd(emb, fc1w, fc1b)
        r_act = _activate(r_pre, ops.ACT_GELU_ERF)  # element-wise pass on the saved pre-activation (no second GEMM)
        distill = ops.eltwise(ops.ELT_AXPY, _lin_fwd(r_act, fc2w, fc2b), emb, scale=cfg["alpha"])
        pooled, _ = ops.mean_rows(emb.view(B, T, D), want32=True)
        c_pre = _lin_fwd(pooled, c1w, c1b)
        c_act = _activate(c_pre, ops.ACT_RELU)
        logits = _lin_fwd(c_act, c2w, c2b)
        ctx.cfg, ctx.saved = cfg, saved
        ctx.save_for_backward(*params)  # version-checked by autograd
        ctx.misc = dict(patches=patches, z0=z0, z_cls=z_cls, cls16=cls16, emb=emb, r_pre=r_pre, r_act=r_act, pooled=pooled,
                        c_pre=c_pre, c_act=c_act)
        return emb.view(B, T, D), distill.view(B, T, D), logits

    @staticmethod
    def backward(ctx, d_emb, d_distill, d_logits):
        cfg, params, m = ctx.cfg, ctx.saved_tensors, ctx.misc
        p, d, heads, n_layers, F_, B, T = cfg["patch"], cfg["d"], cfg["heads"], cfg["layers"], cfg["F"], cfg["B"], cfg["T"]
        g = cfg["res"] // p
        n, L = g * g, g * g + 1
        M = F_ * L
        kp = 3 * p * p
        nb = len(_PER_BLOCK)
        conv_w, cls_emb, pos, gpre, bpre = params[:5]
        blocks = [params[5 + i * nb:5 + (i + 1) * nb] for i in range(n_layers)]
        gpost, bpost, proj = params[5 + n_layers * nb:5 + n_layers * nb + 3]
        fc1w, fc1b, fc2w, fc2b, c1w, c1b, c2w, c2b = params[5 + n_layers * nb + 3:]
        dev = m["emb"].device
        D = m["emb"].shape[1]
        zeros = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)  # noqa: E731
        d_emb = zeros(F_, D) if d_emb is None else d_emb.reshape(F_, D).float().contiguous()
        d_distill = zeros(F_, D) if d_distill is None else d_distill.reshape(F_, D).float().contiguous()
        d_logits = zeros(B, c2w.shape[0]) if d_logits is None else d_logits.float().contiguous()
        # ---- heads ----
        dc_act, g_c2w, g_c2b = _lin_bwd(d_logits, m["c_act"], c2w)
        dc_pre = ops.eltwise(ops.ELT_RELU_BWD, dc_act, m["c_pre"])
        d_pooled, g_c1w, g_c1b = _lin_bwd(dc_pre, m["pooled"], c1w)
        dr_out = ops.eltwise(ops.ELT_SCALE, d_distill, scale=cfg["alpha"])
        dr_act, g_fc2w, g_fc2b = _lin_bwd(dr_out, m["r_act"], fc2w)
        dr_pre = ops.eltwise(ops.ELT_GELU_BWD, dr_act, m["r_pre"])
        d_emb_r, g_fc1w, g_fc1b = _lin_bwd(dr_pre, m["emb"], fc1w)
        de = ops.eltwise(ops.ELT_ADD, d_emb, d_distill)
        de = ops.eltwise(ops.ELT_ADD, de, d_emb_r)
        de = ops.eltwise(ops.ELT_ADD, de, ops.broadcast_rows(d_pooled, T, 1.0 / T))
        # ---- proj and ln_post (CLS rows only) ----
        de16 = ops.cast_bf16(de)
        d_cls = _gemm16(de16, _bf(proj), k=D, out_dtype=torch.float32)  # [F, d] = de [F, D] proj^T: W operand = proj [d, D]
        g_proj = _gemm16(m["cls16"], de16, a_t=True, w_t=True, out_dtype=torch.float32)  # [d, D] = cls^T de
        dz_cls, g_gpost, g_bpost = ops.layernorm_bwd(m["z_cls"], gpost.detach().float(), 1e-5, d_cls)
        dx = zeros(M, d)
        dx.view(F_, L, d)[:, 0, :] = dz_cls
        grads_blocks = [None] * n_layers
        for li in reversed(range(n_layers)):
            (g1, b1_, wqkv, bqkv, wo, bo, g2, b2_, w1, bb1, w2, bb2) = blocks[li]
            s = ctx.saved[li]
            # x_next = x_mid + c_proj(QuickGELU(c_fc(ln_2(x_mid))))
            dh, g_w2, g_b2 = _lin_bwd16(dx, s["h16"], w2)
            dxn2, g_w1, g_b1 = _lin_bwd16(dh, s["xn2"], w1, qgelu_pre=s["h_pre"])  # QuickGELU backward fused into the cast pass
            dx_mid, g_g2, g_bt2 = ops.layernorm_bwd(s["x_mid"], g2.detach().float(), 1e-5, dxn2, add=dx)  # + residual branch
            # x_mid = x + out_proj(attn(ln_1(x)))
            da, g_wo, g_bo = _lin_bwd16(dx_mid, s["a"], wo)
            dqkv = ops.attention_vit_bwd(s["qkv"], da, F_, L, heads)
            dxn, g_wqkv, g_bqkv = _lin_bwd16(dqkv, s["xn"], wqkv)
            dx, g_g1, g_bt1 = ops.layernorm_bwd(s["x"], g1.detach().float(), 1e-5, dxn, add=dx_mid)
            grads_blocks[li] = [g_g1, g_bt1, g_wqkv, g_bqkv, g_wo, g_bo, g_g2, g_bt2, g_w1, g_b1, g_w2, g_b2]
        # ---- ln_pre, embeddings, patch embedding ----
        dz0, g_gpre, g_bpre = ops.layernorm_bwd(m["z0"], gpre.detach().float(), 1e-5, dx)
        g_pos = ops.colsum(dz0.view(F_, L * d)).view(L, d)  # sum over frames
        g_cls = g_pos[0].clone()
        dz_tok = dz0.view(F_, L, d)[:, 1:, :].reshape(F_ * n, d).contiguous()  # token rows (strided copy: plumbing)
        g_conv = _gemm16(ops.cast_bf16(dz_tok), m["patches"][:, :kp], a_t=True, w_t=True, out_dtype=torch.float32)  # dz_tok^T patches
        grads = [g_conv.view_as(conv_w), g_cls, g_pos, g_gpre, g_bpre]
        for gb in grads_blocks:
            grads += gb
        grads += [g_gpost, g_bpost, g_proj, g_fc1w, g_fc1b, g_fc2w, g_fc2b, g_c1w, g_c1b, g_c2w, g_c2b]
        return (None, None, *grads)


def student_train_forward(model, patches: torch.Tensor, B: int, T: int):
    """Training-mode ``encode_patches`` of the student (called by ``_StudentBase.forward`` when ``self.training``)."""
    tower = model.visual_encoder
    _, params = trainable_parameters(model)
    cfg = dict(patch=tower.patch_size, d=tower.width, heads=tower.heads, layers=tower.layers, res=tower.input_resolution,
               F=B * T, B=B, T=T, alpha=float(model.residual_mlp.alpha))
    return StudentTrainFunction.apply(cfg, patches, *params)
