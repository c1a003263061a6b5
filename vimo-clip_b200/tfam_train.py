"""TFAM training step: forward + backward of ``AMO_CLIP`` in ``.train()`` mode on the sm_100a kernels.

Reference: ``TFAM/train_and_eval.py:66-101`` (``ModelTrainer.train_epoch``): ``output = model(...)``,
``loss = BCEWithLogitsLoss(output, labels)``, ``loss.backward()``, ``AdamW.step()``.  The drop-in keeps that
call sequence: in training mode ``AMO_CLIP.forward`` returns logits attached to the autograd graph through
:class:`TfamTrainFunction`, whose ``backward`` produces the gradient of every parameter, so
``loss.backward()``, ``torch.optim.AdamW`` and ``torch.nn.parallel.DistributedDataParallel`` (whose bucketed
NCCL all-reduce hooks fire on the parameters' AccumulateGrad nodes) work unchanged.

Arithmetic (all fusion modes of ``AMO_CLIP.py:136-167``; the default is cross-attention, ``:146-150,170``):
  * every ``nn.Linear`` forward is one tcgen05 GEMM on split-bf16 operands; its backward is two more
    (``dX = dY W`` and ``dW = dY^T X``) fed by ``ops.transpose_split``; bias gradients are column sums;
  * LayerNorm, masked attention (with dropout on the probabilities), ReLU / GELU, dropout and the temporal
    mean have fp32 forward / backward kernels (``csrc/backward.cu``);
  * dropout masks are drawn with ``torch.rand`` on the device (RNG plumbing) and applied by our kernels, so the
    random stream differs from the reference's fused dropout; with ``dropout = mlp_dropout = 0`` the step is
    deterministic and is checked against fp32 autograd of the reference module.
"""
from __future__ import annotations

import torch

from . import ops
from ._params import named_tensors

_PER_LAYER = ("self_attn.in_proj_weight", "self_attn.in_proj_bias", "self_attn.out_proj.weight", "self_attn.out_proj.bias",
              "cross_attn.in_proj_weight", "cross_attn.in_proj_bias", "cross_attn.out_proj.weight", "cross_attn.out_proj.bias",
              "ffn.0.weight", "ffn.0.bias", "ffn.3.weight", "ffn.3.bias",
              "norm_self.weight", "norm_self.bias", "norm_cross.weight", "norm_cross.bias", "norm_ffn.weight", "norm_ffn.bias")
_HEAD = ("classifier.0.weight", "classifier.0.bias", "classifier.1.weight", "classifier.1.bias", "classifier.4.weight", "classifier.4.bias")


def trainable_parameters(model, proj: bool = False):
    """The parameters in the order TfamTrainFunction takes them (projection_layer only in the embedding-concat mode)."""
    named = dict(named_tensors(model))  # also on DataParallel replicas (TFAM/train_and_eval.py:392), whose _parameters are empty
    names = [f"layers.{i}.{n}" for i in range(len(model.layers)) for n in _PER_LAYER] + list(_HEAD)
    if proj:
        names += ["projection_layer.weight", "projection_layer.bias"]
    return names, [named[n] for n in names]


def _lin_fwd(x32, w, b, act=ops.ACT_NONE):
    return ops.gemm(ops.cast_bf16(x32, split=True), ops.split_weight(w), bias=b, act=act, out_dtype=torch.float32)


def _lin_bwd(dy, x32, w, need_dx=True, dw_out=None):
    """y = x W^T + b  ->  (dx | None, dW, db), all fp32, through the split-bf16 tcgen05 GEMM."""
    M, N = dy.shape
    dx = None
    if need_dx:  # dx[M,K] = dy[M,N] W[N,K]: "weight" operand = W^T stored [K, 3N]
        dx = ops.gemm(ops.cast_bf16(dy, split=True), ops.transpose_split(w, 1), k=3 * N, out_dtype=torch.float32)
    # dW[N,K] = sum_m dy[m,n] x[m,k]: A = dy^T [N, 3M], "weight" operand = x^T [K, 3M]
    dw = ops.gemm(ops.transpose_split(dy, 0), ops.transpose_split(x32, 1), k=3 * M, out_dtype=torch.float32, out=dw_out)
    return dx, dw, ops.colsum(dy)


def _activate(pre, act):
    """ReLU / GELU(erf) of a saved pre-activation (one element-wise kernel instead of a second GEMM with the activation epilogue)."""
    if act == ops.ACT_RELU:
        return ops.eltwise(ops.ELT_RELU_BWD, pre, pre)  # pre * (pre > 0)
    return ops.eltwise(ops.ELT_GELU_FWD, pre, None)


def _dropout_mask(shape, p, dev):
    if p <= 0.0:
        return None
    return (torch.rand(shape, device=dev) >= p).to(torch.float32).mul_(1.0 / (1.0 - p)).contiguous()


def _drop(x, mask):
    return x if mask is None else ops.eltwise(ops.ELT_MUL, x, mask.view_as(x))


class TfamTrainFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, x_in, mot, v_rgb, v_mot, *params):
        """x_in [B,T,d] (or [B,T,2d] when cfg["proj"]: the embedding-concat mode projects it first); mot = cross-attention
        source [B,Tm,d] or None (rgb-only / flow-only / concat modes: AttentionLayer.forward skips the cross block)."""
        d, h, p_drop, p_mlp, eps_list = cfg["d"], cfg["heads"], cfg["dropout"], cfg["mlp_dropout"], cfg["eps"]
        n_layers = cfg["layers"]
        cross = mot is not None
        dev = x_in.device
        B, T, _ = x_in.shape
        Tm = mot.shape[1] if cross else 0
        M, Mm = B * T, B * Tm
        k = len(_PER_LAYER)
        x = x_in.reshape(M, x_in.shape[2]).contiguous()
        proj_in = None
        if cfg["proj"]:  # AMO_CLIP.py:163-165: x = projection_layer(cat([rgb[:, :-1], motion], -1))
            wp, bp = params[n_layers * k + len(_HEAD):]
            proj_in = x
            x = _lin_fwd(x, wp, bp)
        m32 = mot.reshape(Mm, d).contiguous() if cross else None
        saved = []
        for li in range(n_layers):
            (w_sin, b_sin, w_so, b_so, w_cin, b_cin, w_co, b_co, w1, b1, w2, b2, gs, bs, gc, bc, gf, bf_) = params[li * k:(li + 1) * k]
            e_s, e_c, e_f = eps_list[li]
            # self-attention block (AMO_CLIP.py:39-40)
            qkv = _lin_fwd(x, w_sin, b_sin)
            pm1 = _dropout_mask((B, h, T, T), p_drop, dev)
            a1 = ops.attention_masked(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], v_rgb, B, T, T, h, out_dtype=torch.float32, prob_mask=pm1)
            dm1 = _dropout_mask((M, d), p_drop, dev)
            z1 = ops.eltwise(ops.ELT_ADD, x, _drop(_lin_fwd(a1, w_so, b_so), dm1))
            x1, _ = ops.layernorm(z1, gs, bs, eps=e_s, want32=True, want16=False)
            # cross-attention block (AMO_CLIP.py:43-45), skipped when there is no cross source
            q2 = kv = pm2 = a2 = dm2 = z2 = None
            x2 = x1
            if cross:
                q2 = _lin_fwd(x1, w_cin[:d], b_cin[:d])
                kv = _lin_fwd(m32, w_cin[d:], b_cin[d:])
                pm2 = _dropout_mask((B, h, T, Tm), p_drop, dev)
                a2 = ops.attention_masked(q2, kv[:, :d], kv[:, d:], v_mot, B, T, Tm, h, out_dtype=torch.float32, prob_mask=pm2)
                dm2 = _dropout_mask((M, d), p_drop, dev)
                z2 = ops.eltwise(ops.ELT_ADD, x1, _drop(_lin_fwd(a2, w_co, b_co), dm2))
                x2, _ = ops.layernorm(z2, gc, bc, eps=e_c, want32=True, want16=False)
            # feed-forward block (AMO_CLIP.py:48-49)
            h_pre = _lin_fwd(x2, w1, b1)  # one GEMM: the backward needs the pre-activation, the activation is an element-wise pass
            h_act = _activate(h_pre, cfg["act"][li])
            dmh = _dropout_mask(tuple(h_act.shape), p_drop, dev)
            h_d = _drop(h_act, dmh)
            dm3 = _dropout_mask((M, d), p_drop, dev)
            z3 = ops.eltwise(ops.ELT_ADD, x2, _drop(_lin_fwd(h_d, w2, b2), dm3))
            x3, _ = ops.layernorm(z3, gf, bf_, eps=e_f, want32=True, want16=False)
            saved.append(dict(x=x, qkv=qkv, a1=a1, z1=z1, x1=x1, q2=q2, kv=kv, a2=a2, z2=z2, x2=x2, h_pre=h_pre, h_d=h_d, z3=z3,
                              pm1=pm1, dm1=dm1, pm2=pm2, dm2=dm2, dmh=dmh, dm3=dm3))
            x = x3
        g0, b0, wc1, bc1, wc2, bc2 = params[n_layers * k:n_layers * k + len(_HEAD)]
        pooled, _ = ops.mean_rows(x.view(B, T, d), want32=True)  # ALL rows, padded ones included (AMO_CLIP.py:170)
        n32, _ = ops.layernorm(pooled, g0, b0, eps=cfg["head_eps"], want32=True, want16=False)
        u_pre = _lin_fwd(n32, wc1, bc1)
        u = _activate(u_pre, ops.ACT_GELU_ERF)
        dmu = _dropout_mask(tuple(u.shape), p_mlp, dev)
        u_d = _drop(u, dmu)
        logits = _lin_fwd(u_d, wc2, bc2)
        ctx.cfg, ctx.saved, ctx.head = cfg, saved, dict(pooled=pooled, n32=n32, u_pre=u_pre, u_d=u_d, dmu=dmu)
        ctx.m32, ctx.v_rgb, ctx.v_mot, ctx.shape = m32, v_rgb, v_mot, (B, T, Tm)
        ctx.proj_in, ctx.cross = proj_in, cross
        ctx.save_for_backward(*params)  # version-checked: an in-place parameter update between forward and backward is an error
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        cfg, params = ctx.cfg, ctx.saved_tensors
        d, h, n_layers = cfg["d"], cfg["heads"], cfg["layers"]
        B, T, Tm = ctx.shape
        M, Mm = B * T, B * Tm
        k = len(_PER_LAYER)
        grads = [None] * len(params)
        hd = ctx.head
        g0, b0, wc1, bc1, wc2, bc2 = params[n_layers * k:n_layers * k + len(_HEAD)]
        cross = ctx.cross
        dlogits = dlogits.float().contiguous()
        du_d, gw2, gb2 = _lin_bwd(dlogits, hd["u_d"], wc2)
        du = _drop(du_d, hd["dmu"])
        du_pre = ops.eltwise(ops.ELT_GELU_BWD, du, hd["u_pre"])
        dn, gw1, gb1 = _lin_bwd(du_pre, hd["n32"], wc1)
        dpooled, gg0, gb0 = ops.layernorm_bwd(hd["pooled"], g0, cfg["head_eps"], dn)
        grads[n_layers * k:n_layers * k + len(_HEAD)] = [gg0, gb0, gw1, gb1, gw2, gb2]
        dx = ops.broadcast_rows(dpooled, T, 1.0 / T)  # gradient w.r.t. the last layer's output rows
        for li in reversed(range(n_layers)):
            (w_sin, b_sin, w_so, b_so, w_cin, b_cin, w_co, b_co, w1, b1, w2, b2, gs, bs, gc, bc, gf, bf_) = params[li * k:(li + 1) * k]
            e_s, e_c, e_f = cfg["eps"][li]
            s = ctx.saved[li]
            # ---- feed-forward block ----
            dz3, g_gf, g_bf = ops.layernorm_bwd(s["z3"], gf, e_f, dx)
            dh_d, g_w2, g_b2 = _lin_bwd(_drop(dz3, s["dm3"]), s["h_d"], w2)
            dh = _drop(dh_d, s["dmh"])
            dh_pre = ops.eltwise(ops.ELT_RELU_BWD if cfg["act"][li] == ops.ACT_RELU else ops.ELT_GELU_BWD, dh, s["h_pre"])
            dx2_b, g_w1, g_b1 = _lin_bwd(dh_pre, s["x2"], w1)
            dx2 = ops.eltwise(ops.ELT_ADD, dz3, dx2_b)
            # ---- cross-attention block ----
            g_wcin = g_bcin = g_wco = g_bco = g_gc = g_bc = None  # unused parameters get no gradient, as in the reference
            dx1 = dx2
            if cross:
                dz2, g_gc, g_bc = ops.layernorm_bwd(s["z2"], gc, e_c, dx2)
                da2, g_wco, g_bco = _lin_bwd(_drop(dz2, s["dm2"]), s["a2"], w_co)
                dq2 = torch.empty((M, d), dtype=torch.float32, device=dx.device)
                dkv = torch.empty((Mm, 2 * d), dtype=torch.float32, device=dx.device)
                ops.attention_masked_bwd(s["q2"], s["kv"][:, :d], s["kv"][:, d:], ctx.v_mot, s["pm2"], da2, B, T, Tm, h,
                                         dq2, dkv[:, :d], dkv[:, d:])
                g_wcin = torch.empty_like(w_cin, dtype=torch.float32)
                dx1_b, _, g_bq = _lin_bwd(dq2, s["x1"], w_cin[:d], dw_out=g_wcin[:d])
                _, _, g_bkv = _lin_bwd(dkv, ctx.m32, w_cin[d:], need_dx=False, dw_out=g_wcin[d:])
                g_bcin = torch.cat([g_bq, g_bkv])
                dx1 = ops.eltwise(ops.ELT_ADD, dz2, dx1_b)
            # ---- self-attention block ----
            dz1, g_gs, g_bs = ops.layernorm_bwd(s["z1"], gs, e_s, dx1)
            da1, g_wso, g_bso = _lin_bwd(_drop(dz1, s["dm1"]), s["a1"], w_so)
            dqkv = torch.empty((M, 3 * d), dtype=torch.float32, device=dx.device)
            qkv = s["qkv"]
            ops.attention_masked_bwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], ctx.v_rgb, s["pm1"], da1, B, T, T, h,
                                     dqkv[:, :d], dqkv[:, d:2 * d], dqkv[:, 2 * d:])
            need_dx = li > 0 or cfg["proj"]
            dx_b, g_wsin, g_bsin = _lin_bwd(dqkv, s["x"], w_sin, need_dx=need_dx)
            if need_dx:
                dx = ops.eltwise(ops.ELT_ADD, dz1, dx_b)
            grads[li * k:(li + 1) * k] = [g_wsin, g_bsin, g_wso, g_bso, g_wcin, g_bcin, g_wco, g_bco,
                                          g_w1, g_b1, g_w2, g_b2, g_gs, g_bs, g_gc, g_bc, g_gf, g_bf]
        if cfg["proj"]:
            wp = params[n_layers * k + len(_HEAD)]
            _, g_wp, g_bp = _lin_bwd(dx, ctx.proj_in, wp, need_dx=False)
            grads[n_layers * k + len(_HEAD):] = [g_wp, g_bp]
        return (None, None, None, None, None, *grads)


def tfam_train_forward(model, x_in, cross_src, v_x, v_cross, proj: bool = False):
    """Training-mode forward of ``AMO_CLIP`` (called by ``AMO_CLIP.forward`` when ``self.training``): ``x_in`` is the
    sequence the layers run on (already selected / concatenated per fusion mode), ``cross_src`` the cross-attention
    source or None."""
    _, params = trainable_parameters(model, proj)
    cfg = dict(
        proj=proj,
        d=model.d_model, heads=model.nhead, layers=len(model.layers),
        dropout=float(model.layers[0].dropout.p), mlp_dropout=float(model.classifier[3].p),
        eps=[(ly.norm_self.eps, ly.norm_cross.eps, ly.norm_ffn.eps) for ly in model.layers],
        act=[ops.ACT_GELU_ERF if ly.activation == "gelu" else ops.ACT_RELU for ly in model.layers],
        head_eps=model.classifier[0].eps,
    )
    return TfamTrainFunction.apply(cfg, x_in, cross_src, v_x, v_cross, *params)
