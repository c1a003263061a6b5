"""Tensor-level wrappers over the C-ABI.  PyTorch is used only for device memory and streams.

Every function enqueues hand-written sm_100a kernels on the current CUDA stream and returns
device tensors.  CPU tensors raise: there is no fallback path.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import ACT_GELU_ERF, ACT_NONE, ACT_QUICKGELU, ACT_RELU  # noqa: F401  (re-exported)


def _need_cuda(*ts) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.VmcError("vimoclip_b200 ops run on CUDA tensors only (no CPU fallback)")


def _p(t) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count() -> int:
    return int(_lib.lib().vmc_launch_count())


def reset_launch_count() -> None:
    _lib.lib().vmc_reset_launch_count()


def set_option(option: int, value: int) -> None:
    """Select an implementation variant (``_lib.OPT_GEMM_IMPL`` / ``_lib.OPT_ATTN_IMPL``); 0 = default."""
    _lib.check(_lib.lib().vmc_set_option(option, value), "vmc_set_option")


def device_info():
    sm, major, minor = C.c_int(), C.c_int(), C.c_int()
    _lib.check(_lib.lib().vmc_device_info(C.byref(sm), C.byref(major), C.byref(minor)), "vmc_device_info")
    return sm.value, major.value, minor.value


# ---------------------------------------------------------------------------------------------
# P1 prologue
# ---------------------------------------------------------------------------------------------
def patch_ld(patch: int) -> int:
    """Row stride of the patchified GEMM operand: 3*p*p rounded up to 8 elements (16 bytes)."""
    k = 3 * patch * patch
    return (k + 7) // 8 * 8


def prologue(frames: torch.Tensor, *, wrap: bool, dst: str, patch: int = 0, normalised: bool = False) -> torch.Tensor:
    """frames [F,3,H,W] uint8 or float32 -> 'u8' | 'f32' | 'patch' output.

    ``wrap=True`` reproduces the student's in-forward ``.float()`` + ``to_pil_image`` wrap
    (models/student_model.py:74,78); ``wrap=False`` takes uint8 pixels as they are (the HF
    ``CLIPImageProcessor`` path of extract_embeddings.py:91).  Float input always wraps.
    """
    _need_cuda(frames)
    if frames.dim() != 4 or frames.shape[1] != 3:
        raise ValueError("frames must be [F,3,H,W]")
    frames = frames.contiguous()
    F_, _, H, W = frames.shape
    if frames.dtype == torch.uint8:
        src_kind = _lib.SRC_U8_WRAP if wrap else _lib.SRC_U8
    elif frames.dtype == torch.float32:
        # normalised=True: the tensor already holds (u8/255 - mean)/std (get_image_features input)
        src_kind = _lib.SRC_F32_NORM if normalised else _lib.SRC_F32_WRAP
    else:
        raise TypeError("frames must be uint8 or float32")
    ld = 0
    if dst == "u8":
        out = torch.empty((F_, 3, H, W), dtype=torch.uint8, device=frames.device)
        kind = _lib.DST_U8
    elif dst == "f32":
        out = torch.empty((F_, 3, H, W), dtype=torch.float32, device=frames.device)
        kind = _lib.DST_F32_NCHW
    elif dst == "patch":
        if patch <= 0 or H % patch or W % patch:
            raise ValueError("H and W must be multiples of the patch size")
        ld = patch_ld(patch)
        out = torch.empty((F_ * (H // patch) * (W // patch), ld), dtype=torch.bfloat16, device=frames.device)
        kind = _lib.DST_BF16_PATCH
    else:
        raise ValueError(dst)
    with torch.cuda.device(frames.device):
        _lib.check(_lib.lib().vmc_prologue(_p(frames), src_kind, _p(out), kind, F_, H, W, patch, ld, _stream()), "vmc_prologue")
    return out


def resize_geometry(H: int, W: int, size: int = 224, hf_crop: bool = False):
    """torchvision Resize(size) + CenterCrop(size) geometry: (new_h, new_w, top, left); hf_crop: HF's floor crop offset."""
    v = [C.c_int() for _ in range(4)]
    _lib.check(_lib.lib().vmc_resize_geometry_ex(H, W, size, 1 if hf_crop else 0, *[C.byref(x) for x in v]), "vmc_resize_geometry")
    return tuple(x.value for x in v)


def resize_center_crop(frames: torch.Tensor, *, wrap: bool, size: int = 224, hf_crop: bool = False) -> torch.Tensor:
    """frames [F,3,H,W] uint8/float32 -> uint8 [F,3,size,size]: (optional student wrap) -> Pillow bicubic resize of the
    short side to ``size`` (up or down) -> centre crop, bit-exact with ``Resize(224, BICUBIC)`` + ``CenterCrop(224)`` on PIL
    images; hf_crop = True takes HF CLIPImageProcessor's crop offset ((dim - size) // 2) instead of torchvision's round."""
    _need_cuda(frames)
    if frames.dim() != 4 or frames.shape[1] != 3:
        raise ValueError("frames must be [F,3,H,W]")
    frames = frames.contiguous()
    F_, _, H, W = frames.shape
    if frames.dtype == torch.uint8:
        src_kind = _lib.SRC_U8_WRAP if wrap else _lib.SRC_U8
    elif frames.dtype == torch.float32:
        src_kind = _lib.SRC_F32_WRAP
    else:
        raise TypeError("frames must be uint8 or float32")
    out = torch.empty((F_, 3, size, size), dtype=torch.uint8, device=frames.device)
    tmp = torch.empty((F_, 3, H, size), dtype=torch.uint8, device=frames.device)
    with torch.cuda.device(frames.device):
        _lib.check(_lib.lib().vmc_resize_center_crop_ex(_p(frames), src_kind, _p(out), _p(tmp), F_, H, W, size, 1 if hf_crop else 0,
                                                        _stream()), "vmc_resize_center_crop")
    return out


def frame_diff(bgr: torch.Tensor, *, dst: str | None = None, patch: int = 0, want_diff: bool = True):
    """bgr [clips, T+1, H, W, 3] uint8 -> (diff_u8 [clips,T,H,W] | None, student prologue output | None).

    utils/generate_frame_diff_video.py:37,46,49 fused with the student prologue of the diff frames.
    """
    _need_cuda(bgr)
    if bgr.dtype != torch.uint8 or bgr.dim() != 5 or bgr.shape[-1] != 3:
        raise ValueError("bgr must be uint8 [clips, T+1, H, W, 3]")
    bgr = bgr.contiguous()
    clips, T1, H, W, _ = bgr.shape
    T = T1 - 1
    if T < 1:
        raise ValueError("need at least two frames per clip")
    diff = torch.empty((clips, T, H, W), dtype=torch.uint8, device=bgr.device) if want_diff else None
    out, kind, ld = None, 0, 0
    Fo = clips * T
    if dst == "u8":
        out, kind = torch.empty((Fo, 3, H, W), dtype=torch.uint8, device=bgr.device), _lib.DST_U8
    elif dst == "f32":
        out, kind = torch.empty((Fo, 3, H, W), dtype=torch.float32, device=bgr.device), _lib.DST_F32_NCHW
    elif dst == "patch":
        ld = patch_ld(patch)
        out = torch.empty((Fo * (H // patch) * (W // patch), ld), dtype=torch.bfloat16, device=bgr.device)
        kind = _lib.DST_BF16_PATCH
    elif dst is not None:
        raise ValueError(dst)
    with torch.cuda.device(bgr.device):
        _lib.check(
            _lib.lib().vmc_frame_diff_prologue(_p(bgr), _p(diff), _p(out), kind, clips, T, H, W, patch, ld, _stream()),
            "vmc_frame_diff_prologue",
        )
    return diff, out


# ---------------------------------------------------------------------------------------------
# GEMM / LayerNorm / attention
# ---------------------------------------------------------------------------------------------
def gemm(a: torch.Tensor, w: torch.Tensor, *, bias=None, act: int = ACT_NONE, alpha: float = 1.0, resid=None,
         out: torch.Tensor | None = None, out_dtype=torch.bfloat16, n: int | None = None, k: int | None = None,
         row_group: int = 0, out_rows: int | None = None, emit_stats=None, fold=None, a_t: bool = False,
         w_t: bool = False) -> torch.Tensor:
    """out = alpha * act(a @ w.T + bias) + resid.  a [M, lda] bf16, w [N, ldw] bf16 (nn.Linear layout).
    resid may be fp32 or bf16 (bf16 residual stream of the ViT tower).
    emit_stats = (raw16 bf16 [M, N] | None, stats fp32 [parts, M, 2]): producer side of a folded LayerNorm; raw16 None with
    a bf16 resid and a bf16 out: the output IS the bf16 row copy and the statistics are those of the rounded values.
    fold = (stats fp32 [parts, M, 2], colsum fp32 [N], eps): consumer side (a = raw rows, w = gamma-folded weights).
    a_t / w_t: the operand is given transposed in memory (a = A^T [K, M], w = W^T [K, N], row-major): MN-major UMMA
    operands, no transposing pass (the dW = dY^T X and dX = dY W GEMMs of the backward)."""
    _need_cuda(a, w, bias, resid, out)
    if a.dtype != torch.bfloat16 or w.dtype != torch.bfloat16:
        raise TypeError("gemm operands must be bf16")
    if a.stride(-1) != 1 or w.stride(-1) != 1:
        raise ValueError("gemm operands must be row-major")
    M = a.shape[1] if a_t else a.shape[0]
    K = (a.shape[0] if a_t else a.shape[1]) if k is None else k
    N = (w.shape[1] if w_t else w.shape[0]) if n is None else n
    if out is None:
        rows = M if out_rows is None else out_rows
        out = torch.empty((rows, N), dtype=out_dtype, device=a.device)
    e = _lib.GemmEpilogue()
    e.bias = 0 if bias is None else bias.data_ptr()
    e.resid = 0 if resid is None else resid.data_ptr()
    e.ldr = 0 if resid is None else resid.stride(0)
    e.out = out.data_ptr()
    e.ldo = out.stride(0)
    e.out_bf16 = 1 if out.dtype == torch.bfloat16 else 0
    if out.dtype not in (torch.bfloat16, torch.float32):
        raise TypeError("gemm output must be bf16 or fp32")
    if bias is not None and bias.dtype != torch.float32:
        raise TypeError("bias must be fp32")
    if resid is not None and resid.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError("resid must be fp32 or bf16")
    e.resid_bf16 = 1 if (resid is not None and resid.dtype == torch.bfloat16) else 0
    e.act = act
    e.alpha = alpha
    e.row_group = row_group
    if emit_stats is not None:
        raw16, stats = emit_stats
        _need_cuda(raw16, stats)
        if raw16 is not None:
            e.raw16_out, e.raw16_ld = raw16.data_ptr(), raw16.stride(0)
        e.stats_out, e.stats_ld = stats.data_ptr(), stats.shape[1]
    if fold is not None:
        stats, colsum, eps_ = fold
        _need_cuda(stats, colsum)
        e.stats_in, e.stats_parts, e.stats_ld, e.colsum, e.ln_eps = stats.data_ptr(), stats.shape[0], stats.shape[1], colsum.data_ptr(), eps_
    with torch.cuda.device(a.device):
        _lib.check(_lib.lib().vmc_gemm_bf16_ex(_p(a), a.stride(0), 1 if a_t else 0, _p(w), w.stride(0), 1 if w_t else 0, M, N, K,
                                               C.byref(e), _stream()), "vmc_gemm_bf16")
    return out


def gemm_stats_parts(M: int, N: int) -> int:
    """Column slices per row that ``gemm(..., emit_stats=...)`` writes for this shape."""
    return int(_lib.lib().vmc_gemm_stats_parts(M, N))


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, *, eps: float = 1e-5, want32: bool = False,
              want16: bool = True, out32: torch.Tensor | None = None, split16: bool = False, stats: torch.Tensor | None = None,
              stats_rounded: bool = False, cls_row: torch.Tensor | None = None, cls_every: int = 0):
    """x fp32 or bf16 [rows, d] -> (y32 | None, y16 | None).  split16: y16 is the [rows, 3d] split operand [hi|lo|hi].
    stats fp32 [rows, 2]: (sum, sum of squares) of the output row (of its bf16 rounding when stats_rounded).
    cls_row fp32 [d] with cls_every > 0: rows r % cls_every == 0 take their input from cls_row (the CLS token under ln_pre)."""
    _need_cuda(x, gamma, beta, out32, stats, cls_row)
    if x.dtype not in (torch.float32, torch.bfloat16) or x.dim() != 2 or x.stride(1) != 1:
        raise ValueError("layernorm input must be fp32 or bf16 [rows, d] row-major")
    rows, d = x.shape
    y32 = out32 if out32 is not None else (torch.empty((rows, d), dtype=torch.float32, device=x.device) if want32 else None)
    y16 = torch.empty((rows, 3 * d if split16 else d), dtype=torch.bfloat16, device=x.device) if want16 else None
    with torch.cuda.device(x.device):
        _lib.check(
            _lib.lib().vmc_layernorm_ex(_p(x), 1 if x.dtype == torch.bfloat16 else 0, x.stride(0), _p(gamma), _p(beta), eps, _p(y32),
                                        0 if y32 is None else y32.stride(0), _p(y16), 0 if y16 is None else y16.stride(0),
                                        1 if split16 else 0, rows, d, _p(cls_row), cls_every, _p(stats), 1 if stats_rounded else 0, _stream()),
            "vmc_layernorm",
        )
    return y32, y16


def attention_vit(qkv: torch.Tensor, F_: int, L: int, heads: int, impl: int = 5) -> torch.Tensor:
    """qkv bf16 [F*L, 3*heads*64] -> bf16 [F*L, heads*64]; softmax(q k^T / 8) v per (frame, head)."""
    _need_cuda(qkv)
    d = heads * 64
    if qkv.dtype != torch.bfloat16 or tuple(qkv.shape) != (F_ * L, 3 * d) or not qkv.is_contiguous():
        raise ValueError("qkv must be contiguous bf16 [F*L, 3*d]")
    out = torch.empty((F_ * L, d), dtype=torch.bfloat16, device=qkv.device)
    with torch.cuda.device(qkv.device):
        _lib.check(_lib.lib().vmc_attention_vit_impl(_p(qkv), _p(out), F_, L, heads, impl, _stream()), "vmc_attention_vit")
    return out


def attention_cls(q_cls: torch.Tensor, kv: torch.Tensor, F_: int, L: int, heads: int) -> torch.Tensor:
    """q_cls bf16 [F, heads*64] (the CLS query rows), kv bf16 [F*L, 2*heads*64] = [k | v] of every token ->
    bf16 [F, heads*64]: softmax(q k^T / 8) v of the CLS row only (last block of the tower)."""
    _need_cuda(q_cls, kv)
    d = heads * 64
    if q_cls.dtype != torch.bfloat16 or kv.dtype != torch.bfloat16 or tuple(q_cls.shape) != (F_, d) or tuple(kv.shape) != (F_ * L, 2 * d) \
            or not q_cls.is_contiguous() or not kv.is_contiguous():
        raise ValueError("attention_cls: q_cls must be contiguous bf16 [F, d] and kv contiguous bf16 [F*L, 2*d]")
    out = torch.empty((F_, d), dtype=torch.bfloat16, device=q_cls.device)
    with torch.cuda.device(q_cls.device):
        _lib.check(_lib.lib().vmc_attention_cls(_p(q_cls), _p(kv), _p(out), F_, L, heads, _stream()), "vmc_attention_cls")
    return out


def attention_masked(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, key_valid, B: int, Tq: int, Tk: int, heads: int,
                     out_dtype=torch.bfloat16, prob_mask=None) -> torch.Tensor:
    """q [B*Tq, *], k/v [B*Tk, *] fp32 views (row-major, heads*64 columns) -> bf16 or fp32 [B*Tq, heads*64].
    prob_mask fp32 [B, heads, Tq, Tk] (0 or 1/(1-p)): dropout on the attention probabilities (training)."""
    _need_cuda(q, k, v, key_valid, prob_mask)
    if prob_mask is not None and (prob_mask.dtype != torch.float32 or tuple(prob_mask.shape) != (B, heads, Tq, Tk) or not prob_mask.is_contiguous()):
        raise ValueError("prob_mask must be contiguous fp32 [B, heads, Tq, Tk]")
    for t in (q, k, v):
        if t.dtype != torch.float32 or t.stride(1) != 1:
            raise ValueError("attention_masked inputs must be fp32 row-major")
    if key_valid is not None:
        if key_valid.dtype not in (torch.bool, torch.uint8) or tuple(key_valid.shape) != (B, Tk) or not key_valid.is_contiguous():
            raise TypeError("key_valid must be a contiguous bool/uint8 [B, Tk] tensor (True = real frame)")
    if out_dtype not in (torch.bfloat16, torch.float32):
        raise TypeError("attention_masked output must be bf16 or fp32")
    out = torch.empty((B * Tq, heads * 64), dtype=out_dtype, device=q.device)
    with torch.cuda.device(q.device):
        _lib.check(
            _lib.lib().vmc_attention_masked_train(_p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(key_valid),
                                                  _p(prob_mask), _p(out), 1 if out_dtype == torch.float32 else 0, out.stride(0),
                                                  B, Tq, Tk, heads, _stream()),
            "vmc_attention_masked",
        )
    return out


def cast_bf16(x: torch.Tensor, split: bool = False) -> torch.Tensor:
    """fp32 [rows, d] -> bf16 [rows, d], or the split operand [rows, 3d] = [hi | lo | hi] (see ``split_weight``)."""
    _need_cuda(x)
    if x.dtype != torch.float32 or x.dim() != 2 or x.stride(1) != 1:
        raise ValueError("cast_bf16 input must be fp32 [rows, d] row-major")
    rows, d = x.shape
    cols = 3 * d if split else d
    if cols % 8:  # GEMM operands need a row stride that is a multiple of 8 elements: zero-padded tail, pass k=3*d
        y = torch.zeros((rows, (cols + 7) // 8 * 8), dtype=torch.bfloat16, device=x.device)
    else:
        y = torch.empty((rows, cols), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().vmc_cast_bf16(_p(x), x.stride(0), _p(y), y.stride(0), rows, d, 1 if split else 0, _stream()), "vmc_cast_bf16")
    return y


def split_weight(w: torch.Tensor) -> torch.Tensor:
    """One-time packing of an fp32 nn.Linear weight [N, K] as bf16 [N, 3K] = [Whi | Whi | Wlo] so that
    ``gemm(cast_bf16(x, split=True), split_weight(w))`` ~= x @ w.T to ~2^-16 relative (three bf16 MMAs in one
    K = 3K GEMM, fp32 accumulation in TMEM).  Weight packing happens at load time, not on the hot path."""
    w = w.detach().float()
    hi = w.to(torch.bfloat16)
    lo = (w - hi.float()).to(torch.bfloat16)
    return torch.cat([hi, hi, lo], dim=1).contiguous()


def mean_rows(x: torch.Tensor, *, want32: bool = True, want16: bool = False):
    """x fp32 [B, T, d] contiguous -> mean over T (ALL rows, padded ones included)."""
    _need_cuda(x)
    if x.dtype != torch.float32 or x.dim() != 3 or not x.is_contiguous():
        raise ValueError("mean_rows input must be contiguous fp32 [B, T, d]")
    B, T, d = x.shape
    y32 = torch.empty((B, d), dtype=torch.float32, device=x.device) if want32 else None
    y16 = torch.empty((B, d), dtype=torch.bfloat16, device=x.device) if want16 else None
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().vmc_mean_rows(_p(x), _p(y32), _p(y16), B, T, d, _stream()), "vmc_mean_rows")
    return y32, y16


def cosine_distill_loss(student: torch.Tensor, teacher: torch.Tensor) -> torch.Tensor:
    """losses.py:27-40: mean(1 - clamp(cos)) over rows, fp32 scalar on device."""
    _need_cuda(student, teacher)
    s = student.reshape(-1, student.shape[-1]).contiguous().float()
    t = teacher.reshape(-1, teacher.shape[-1]).contiguous().float()
    if s.shape != t.shape:
        raise ValueError("student / teacher shapes differ")
    out = torch.empty((1,), dtype=torch.float32, device=s.device)
    with torch.cuda.device(s.device):
        _lib.check(_lib.lib().vmc_cosine_distill_loss(_p(s), _p(t), s.shape[0], s.shape[1], _p(out), _stream()), "vmc_cosine_distill_loss")
    return out[0]


def student_heads(emb: torch.Tensor, w, alpha: float, num_classes: int):
    """emb fp32 [B,T,D]; w = (fc1_t, b1, fc2_t, b2, c1_t, bc1, c2_t, bc2) with transposed fp32 weights [K,N].
    Returns (distill [B,T,D], logits [B,C]) from ONE fused kernel, one CTA per clip."""
    _need_cuda(emb, *w)
    if emb.dtype != torch.float32 or emb.dim() != 3 or not emb.is_contiguous():
        raise ValueError("emb must be contiguous fp32 [B,T,D]")
    B, T, D = emb.shape
    H = w[4].shape[1]
    distill = torch.empty_like(emb)
    logits = torch.empty((B, num_classes), dtype=torch.float32, device=emb.device)
    with torch.cuda.device(emb.device):
        _lib.check(_lib.lib().vmc_student_heads(_p(emb), _p(w[0]), _p(w[1]), _p(w[2]), _p(w[3]), float(alpha), _p(w[4]), _p(w[5]),
                                                _p(w[6]), _p(w[7]), _p(distill), _p(logits), B, T, D, H, num_classes, _stream()),
                   "vmc_student_heads")
    return distill, logits


def tfam_head(x: torch.Tensor, ln_g, ln_b, eps: float, w1_t, b1, w2_t, b2) -> torch.Tensor:
    """x fp32 [B,T,D] -> logits [B,C]: mean over ALL T rows -> LayerNorm -> Linear -> GELU(erf) -> Linear, fused per clip."""
    _need_cuda(x, ln_g, ln_b, w1_t, b1, w2_t, b2)
    if x.dtype != torch.float32 or x.dim() != 3 or not x.is_contiguous():
        raise ValueError("x must be contiguous fp32 [B,T,D]")
    B, T, D = x.shape
    H, Cn = w1_t.shape[1], w2_t.shape[1]
    logits = torch.empty((B, Cn), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().vmc_tfam_head(_p(x), _p(ln_g), _p(ln_b), float(eps), _p(w1_t), _p(b1), _p(w2_t), _p(b2), _p(logits),
                                            B, T, D, H, Cn, _stream()), "vmc_tfam_head")
    return logits


# ---- backward-pass ops of the TFAM training step (csrc/backward.cu) ----
ELT_MUL, ELT_RELU_BWD, ELT_GELU_BWD, ELT_ADD, ELT_SCALE, ELT_QGELU_FWD, ELT_QGELU_BWD, ELT_AXPY, ELT_GELU_FWD = 0, 1, 2, 3, 4, 5, 6, 7, 8


def _f32_2d(*ts):
    for t in ts:
        if t is not None and (t.dtype != torch.float32 or t.dim() != 2 or t.stride(1) != 1):
            raise ValueError("expected fp32 row-major 2-D tensors")


def transpose_split(x: torch.Tensor, form: int) -> torch.Tensor:
    """x fp32 [R, C] -> bf16 [C, ceil8(3R)]: split operand of x^T (form 0 = [hi|lo|hi], 1 = [hi|hi|lo]); use ``k=3*R``."""
    _need_cuda(x)
    _f32_2d(x)
    R, Cn = x.shape
    ld = (3 * R + 7) // 8 * 8
    y = torch.zeros((Cn, ld), dtype=torch.bfloat16, device=x.device) if ld != 3 * R else torch.empty((Cn, ld), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().vmc_transpose_split(_p(x), x.stride(0), _p(y), y.stride(0), R, Cn, form, _stream()), "vmc_transpose_split")
    return y


def transpose_cast(x: torch.Tensor) -> torch.Tensor:
    """x [R, C] fp32 or bf16 (row-major) -> bf16 [C, ceil8(R)] = x^T (zero-padded tail): pass ``k=R`` to ``gemm``."""
    _need_cuda(x)
    if x.dim() != 2 or x.stride(1) != 1 or x.dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("transpose_cast input must be a row-major fp32 / bf16 matrix")
    R, Cn = x.shape
    ld = (R + 7) // 8 * 8
    y = torch.zeros((Cn, ld), dtype=torch.bfloat16, device=x.device) if ld != R else torch.empty((Cn, ld), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().vmc_transpose_cast(_p(x), 1 if x.dtype == torch.bfloat16 else 0, x.stride(0), _p(y), y.stride(0), R, Cn, _stream()),
                   "vmc_transpose_cast")
    return y


def cast_f32(x: torch.Tensor) -> torch.Tensor:
    """bf16 [rows, d] (row-major view) -> contiguous fp32."""
    _need_cuda(x)
    if x.dim() != 2 or x.stride(1) != 1 or x.dtype != torch.bfloat16:
        raise ValueError("cast_f32 input must be a row-major bf16 matrix")
    y = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().vmc_cast_f32(_p(x), x.stride(0), _p(y), y.stride(0), x.shape[0], x.shape[1], _stream()), "vmc_cast_f32")
    return y


def colsum(x: torch.Tensor, y: torch.Tensor | None = None, out: torch.Tensor | None = None, accumulate: bool = False) -> torch.Tensor:
    """out[c] (+)= sum_r x[r,c] * (y[r,c] if y is given else 1)."""
    _need_cuda(x, y, out)
    _f32_2d(x, y)
    R, Cn = x.shape
    if out is None:
        out = torch.empty(Cn, dtype=torch.float32, device=x.device)
    slices = int(_lib.lib().vmc_colsum_slices(R, Cn))
    ws = torch.empty(slices * Cn, dtype=torch.float32, device=x.device) if slices > 1 else None
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().vmc_colsum(_p(x), x.stride(0), _p(y), 0 if y is None else y.stride(0), _p(out), R, Cn, 1 if accumulate else 0,
                                         _p(ws), _stream()), "vmc_colsum")
    return out


def cast_colsum(x: torch.Tensor, aux: torch.Tensor | None = None):
    """One pass over an fp32 gradient [R, C] (C a multiple of 4): -> (bf16 copy [R, C], column sums [C] of the unrounded values);
    with ``aux`` (the c_fc pre-activation) the QuickGELU backward ``x * qgelu'(aux)`` is applied first."""
    _need_cuda(x, aux)
    _f32_2d(x, aux)
    R, Cn = x.shape
    y = torch.empty((R, Cn), dtype=torch.bfloat16, device=x.device)
    out = torch.empty(Cn, dtype=torch.float32, device=x.device)
    slices = int(_lib.lib().vmc_cast_colsum_slices(R, Cn))
    ws = torch.empty(max(1, slices) * Cn, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().vmc_cast_colsum(_p(x), x.stride(0), _p(aux), 0 if aux is None else aux.stride(0), _p(y), y.stride(0), _p(out),
                                              R, Cn, _p(ws), _stream()), "vmc_cast_colsum")
    return y, out


def layernorm_bwd(z: torch.Tensor, gamma: torch.Tensor, eps: float, dy: torch.Tensor, add: torch.Tensor | None = None):
    """-> (dz [+ add], dgamma, dbeta) for y = LayerNorm(z) * gamma + beta; z is the saved LayerNorm input; ``add`` = the gradient
    that reaches z over the residual branch (pre-LN blocks), summed in the same pass."""
    _need_cuda(z, gamma, dy, add)
    _f32_2d(z, dy, add)
    rows, d = z.shape
    dz = torch.empty_like(z)
    if d in (512, 768, 1024) and gamma.dtype == torch.float32 and gamma.is_contiguous():
        L = _lib.lib()
        blocks = int(L.vmc_layernorm_bwd_fused_blocks(rows))
        ws = torch.empty(blocks * 2 * d, dtype=torch.float32, device=z.device)
        gb = torch.empty(2 * d, dtype=torch.float32, device=z.device)
        with torch.cuda.device(z.device):
            _lib.check(L.vmc_layernorm_bwd_fused(_p(z), z.stride(0), _p(gamma), float(eps), _p(dy), dy.stride(0), _p(add),
                                                 0 if add is None else add.stride(0), _p(dz), dz.stride(0), _p(gb), _p(ws), rows, d, _stream()),
                       "vmc_layernorm_bwd_fused")
        return dz, gb[:d], gb[d:]
    if add is not None:
        dz0, dg, db = layernorm_bwd(z, gamma, eps, dy)
        return eltwise(ELT_ADD, add.contiguous(), dz0), dg, db
    xhat = torch.empty_like(z)
    with torch.cuda.device(z.device):
        _lib.check(_lib.lib().vmc_layernorm_bwd(_p(z), z.stride(0), _p(gamma), float(eps), _p(dy), dy.stride(0), _p(dz), dz.stride(0),
                                                _p(xhat), xhat.stride(0), rows, d, _stream()), "vmc_layernorm_bwd")
    return dz, colsum(dy, xhat), colsum(dy)


def eltwise(mode: int, a: torch.Tensor, b: torch.Tensor | None = None, scale: float = 1.0) -> torch.Tensor:
    _need_cuda(a, b)
    if a.dtype != torch.float32 or not a.is_contiguous() or (b is not None and (b.dtype != torch.float32 or not b.is_contiguous() or b.numel() != a.numel())):
        raise ValueError("eltwise operands must be contiguous fp32 of equal size")
    out = torch.empty_like(a)
    with torch.cuda.device(a.device):
        _lib.check(_lib.lib().vmc_eltwise(mode, _p(a), _p(b), float(scale), _p(out), a.numel(), _stream()), "vmc_eltwise")
    return out


def broadcast_rows(g: torch.Tensor, T: int, scale: float) -> torch.Tensor:
    """g fp32 [B, d] -> [B*T, d] with out[b*T + t] = g[b] * scale (backward of the temporal mean)."""
    _need_cuda(g)
    _f32_2d(g)
    B, d = g.shape
    out = torch.empty((B * T, d), dtype=torch.float32, device=g.device)
    with torch.cuda.device(g.device):
        _lib.check(_lib.lib().vmc_broadcast_rows(_p(g.contiguous()), _p(out), B, T, d, float(scale), _stream()), "vmc_broadcast_rows")
    return out


def attention_masked_bwd(q, k, v, key_valid, prob_mask, dO, B: int, Tq: int, Tk: int, heads: int, dq, dk, dv, tiled=None) -> None:
    """Backward of ``attention_masked``: writes dq [B*Tq, .], dk / dv [B*Tk, .] (fp32 row-major views, heads*64 columns).
    ``tiled``: None = the shared-memory kernel when Tq x Tk fits (~128 x 128), else the tiled two-pass kernels; True forces them."""
    _need_cuda(q, k, v, key_valid, prob_mask, dO, dq, dk, dv)
    _f32_2d(q, k, v, dO, dq, dk, dv)
    need = (2 * Tq * 64 + 2 * Tk * 65 + Tq * (Tk + 1)) * 4
    ws = torch.empty(2 * B * heads * Tq, dtype=torch.float32, device=q.device) if (tiled or need > 200 * 1024) else None
    with torch.cuda.device(q.device):
        _lib.check(
            _lib.lib().vmc_attention_masked_bwd(_p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(key_valid), _p(prob_mask),
                                                _p(dO), dO.stride(0), _p(dq), dq.stride(0), _p(dk), dk.stride(0), _p(dv), dv.stride(0),
                                                B, Tq, Tk, heads, _p(ws), _stream()),
            "vmc_attention_masked_bwd",
        )


def qgelu_cast(x: torch.Tensor) -> torch.Tensor:
    """fp32 (contiguous, numel % 4 == 0) -> bf16 QuickGELU(x) = x * sigmoid(1.702 x)."""
    _need_cuda(x)
    if x.dtype != torch.float32 or not x.is_contiguous() or x.numel() % 4:
        raise ValueError("qgelu_cast input must be contiguous fp32 with numel % 4 == 0")
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().vmc_qgelu_cast(_p(x), _p(y), x.numel(), _stream()), "vmc_qgelu_cast")
    return y


def attention_vit_bwd(qkv: torch.Tensor, dO: torch.Tensor, F_: int, L: int, heads: int) -> torch.Tensor:
    """Backward of ``attention_vit``: qkv bf16 [F*L, 3d] (saved by the forward), dO fp32 [F*L, d] -> dqkv fp32 [F*L, 3d]."""
    _need_cuda(qkv, dO)
    d = heads * 64
    dqkv = torch.empty((F_ * L, 3 * d), dtype=torch.float32, device=qkv.device)
    if L <= 64:
        with torch.cuda.device(qkv.device):
            _lib.check(_lib.lib().vmc_attention_vit_bwd_short(_p(qkv), _p(dO), dO.stride(0), _p(dqkv), F_, L, heads, _stream()),
                       "vmc_attention_vit_bwd_short")
    else:
        q32 = cast_f32(qkv)
        attention_masked_bwd(q32[:, :d], q32[:, d:2 * d], q32[:, 2 * d:], None, None, dO, F_, L, L, heads,
                             dqkv[:, :d], dqkv[:, d:2 * d], dqkv[:, 2 * d:])
    return dqkv
