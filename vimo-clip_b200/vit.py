"""CLIP vision tower on hand-written sm_100a kernels.

``VisionTower`` holds fp32 parameters under the OpenAI ``clip.model.VisionTransformer`` names
(``conv1.weight``, ``class_embedding``, ``positional_embedding``, ``ln_pre``,
``transformer.resblocks.{i}.{ln_1,attn.in_proj_weight,attn.out_proj,ln_2,mlp.c_fc,mlp.c_proj}``,
``ln_post``, ``proj``) so checkpoints written by the reference's ``train.py:167`` load with
``strict=True``.  The forward pass packs those weights once into bf16 [N, K] GEMM operands and runs
``vmc_vit_forward`` (C++ orchestration of the tcgen05 GEMM / attention / LayerNorm kernels).

Reference: ``self.visual_encoder(x)`` at ``models/student_model.py:84``;
``clip_model.get_image_features(pixel_values)`` at ``extract_embeddings.py:94``.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib, ops
from ._params import SharedCache

# name -> (patch, width, layers, heads, output_dim); 224x224 input
VIT_CONFIGS = {
    "ViT-B/32": (32, 768, 12, 12, 512),
    "ViT-B/16": (16, 768, 12, 12, 512),
    "ViT-L/14": (14, 1024, 24, 16, 768),
}


class _Holder(nn.Module):
    """Parameter container (no forward): keeps reference-compatible state_dict key names."""


def _linear_holder(out_f: int, in_f: int) -> _Holder:
    h = _Holder()
    h.weight = nn.Parameter(torch.empty(out_f, in_f))
    h.bias = nn.Parameter(torch.zeros(out_f))
    nn.init.normal_(h.weight, std=in_f**-0.5)
    return h


def _ln_holder(d: int) -> _Holder:
    h = _Holder()
    h.weight = nn.Parameter(torch.ones(d))
    h.bias = nn.Parameter(torch.zeros(d))
    return h


class _Block(_Holder):
    def __init__(self, d: int):
        super().__init__()
        self.attn = _Holder()
        self.attn.in_proj_weight = nn.Parameter(torch.empty(3 * d, d))
        self.attn.in_proj_bias = nn.Parameter(torch.zeros(3 * d))
        nn.init.normal_(self.attn.in_proj_weight, std=d**-0.5)
        self.attn.out_proj = _linear_holder(d, d)
        self.ln_1 = _ln_holder(d)
        self.mlp = _Holder()
        self.mlp.c_fc = _linear_holder(4 * d, d)
        self.mlp.c_proj = _linear_holder(d, 4 * d)
        self.ln_2 = _ln_holder(d)


class VisionTower(nn.Module):
    def __init__(self, patch: int, width: int, layers: int, heads: int, output_dim: int, input_resolution: int = 224,
                 frames_in_flight: int = 2048):
        super().__init__()
        if width != heads * 64:
            raise ValueError("head_dim must be 64 (all CLIP ViT towers)")
        self.input_resolution = input_resolution
        self.patch_size = patch
        self.width, self.layers, self.heads = width, layers, heads
        self.output_dim = output_dim  # read by the student (models/student_model.py:49)
        self.frames_in_flight = frames_in_flight
        g = input_resolution // patch
        self.tokens = g * g + 1
        scale = width**-0.5
        self.conv1 = _Holder()
        self.conv1.weight = nn.Parameter(torch.randn(width, 3, patch, patch) * (3 * patch * patch) ** -0.5)
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn(self.tokens, width))
        self.ln_pre = _ln_holder(width)
        self.transformer = _Holder()
        self.transformer.resblocks = nn.ModuleList([_Block(width) for _ in range(layers)])
        self.ln_post = _ln_holder(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))
        self._cache = SharedCache(self)  # packed weights / workspace per device, shared with DataParallel replicas
        # per-model kernel variants (0 = library default; vmc_vit_model in include/vimoclip_b200.h):
        # ln_mode 6 = bf16 residual stream + folded LayerNorms (default), 3 / 5 = fp32 stream with folds, 4 = fp32 stream with
        # separate LayerNorm kernels; last_block_cls 1 = last block on the CLS rows only (default), 2 = full last block
        self.ln_mode = 0
        self.last_block_cls = 0
        self.attn_impl = 0

    @classmethod
    def from_name(cls, name: str, **kw) -> "VisionTower":
        if name not in VIT_CONFIGS:
            raise ValueError(f"unsupported CLIP vision tower {name!r}; known: {sorted(VIT_CONFIGS)}")
        return cls(*VIT_CONFIGS[name], **kw)

    # ---- weight packing: fp32 state_dict -> bf16 K-major GEMM operands + fp32 vectors ----
    @torch.no_grad()
    def _pack(self):
        dev = self.proj.device
        if dev.type != "cuda":
            raise _lib.VmcError("VisionTower runs on CUDA only (no CPU fallback); call .to('cuda') first")
        self._cache.bind(self)
        return self._cache.get(("pack", dev.index), self._cache.signature(self), self._build_pack)

    @torch.no_grad()
    def _build_pack(self):
        dev = self.proj.device
        keep = []  # tensors referenced by raw pointers below

        def bf(w):
            t = w.detach().to(torch.bfloat16).contiguous()
            keep.append(t)
            return t.data_ptr()

        def f32(w):
            t = w.detach().float().contiguous()
            keep.append(t)
            return t.data_ptr()

        d, p = self.width, self.patch_size
        kp = 3 * p * p
        ld = ops.patch_ld(p)
        wp = torch.zeros((d, ld), dtype=torch.bfloat16, device=dev)
        wp[:, :kp] = self.conv1.weight.detach().reshape(d, kp).to(torch.bfloat16)
        keep.append(wp)
        layers = (_lib.VitLayer * self.layers)()
        for i, blk in enumerate(self.transformer.resblocks):
            ly = layers[i]
            ly.ln1_g, ly.ln1_b = f32(blk.ln_1.weight), f32(blk.ln_1.bias)
            ly.ln2_g, ly.ln2_b = f32(blk.ln_2.weight), f32(blk.ln_2.bias)
            ly.w_qkv, ly.b_qkv = bf(blk.attn.in_proj_weight), f32(blk.attn.in_proj_bias)
            ly.w_out, ly.b_out = bf(blk.attn.out_proj.weight), f32(blk.attn.out_proj.bias)
            ly.w_fc1, ly.b_fc1 = bf(blk.mlp.c_fc.weight), f32(blk.mlp.c_fc.bias)
            ly.w_fc2, ly.b_fc2 = bf(blk.mlp.c_proj.weight), f32(blk.mlp.c_proj.bias)
            # LayerNorm folding (include/vimoclip_b200.h, vmc_gemm_epilogue): LN(x) W^T + b = rstd (x W'^T - mean colsum) + b'
            for ln, lin, names in ((blk.ln_1, (blk.attn.in_proj_weight, blk.attn.in_proj_bias), ("w_qkv_f", "b_qkv_f", "cs_qkv")),
                                   (blk.ln_2, (blk.mlp.c_fc.weight, blk.mlp.c_fc.bias), ("w_fc1_f", "b_fc1_f", "cs_fc1"))):
                w32 = lin[0].detach().float()
                wf = (w32 * ln.weight.detach().float()[None, :]).to(torch.bfloat16).contiguous()
                keep.append(wf)
                setattr(ly, names[0], wf.data_ptr())
                setattr(ly, names[1], f32(lin[1].detach().float() + w32 @ ln.bias.detach().float()))
                setattr(ly, names[2], f32(wf.float().sum(dim=1)))  # of the ROUNDED weights: the mean term cancels exactly
        m = _lib.VitModel()
        m.image, m.patch, m.width, m.layers, m.heads, m.out_dim = self.input_resolution, p, d, self.layers, self.heads, self.output_dim
        m.ld_patch = ld
        m.w_patch = wp.data_ptr()
        m.cls_pos0 = f32(self.class_embedding.detach() + self.positional_embedding.detach()[0])
        m.pos = f32(self.positional_embedding)
        m.ln_pre_g, m.ln_pre_b = f32(self.ln_pre.weight), f32(self.ln_pre.bias)
        m.ln_post_g, m.ln_post_b = f32(self.ln_post.weight), f32(self.ln_post.bias)
        m.w_proj = bf(self.proj.detach().t())
        m.layer = C.cast(layers, C.POINTER(_lib.VitLayer))
        return (m, layers, keep)

    def _get_workspace(self, nbytes: int, device) -> torch.Tensor:
        ws = self._cache.peek(("workspace", device.index))
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
            self._cache.put(("workspace", device.index), ws)
        return ws

    @torch.no_grad()
    def forward_patches(self, patches: torch.Tensor, n_frames: int) -> torch.Tensor:
        """patches bf16 [F*n, ld_patch] (from ``ops.prologue(..., dst='patch')``) -> fp32 [F, output_dim]."""
        m, _, _ = self._pack()
        m.ln_mode, m.last_block_cls, m.attn_impl = int(self.ln_mode), int(self.last_block_cls), int(self.attn_impl)
        n = self.tokens - 1
        out = torch.empty((n_frames, self.output_dim), dtype=torch.float32, device=patches.device)
        L = _lib.lib()
        chunk = max(1, min(self.frames_in_flight, n_frames))
        with torch.cuda.device(patches.device):
            nbytes = int(L.vmc_vit_workspace_bytes(C.byref(m), chunk))
            ws = self._get_workspace(nbytes, patches.device)
            base = (ws.data_ptr() + 1023) // 1024 * 1024
            stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            for f0 in range(0, n_frames, chunk):
                fc = min(chunk, n_frames - f0)
                _lib.check(
                    L.vmc_vit_forward(C.byref(m), C.c_void_p(patches[f0 * n].data_ptr()), C.c_void_p(out[f0].data_ptr()), fc,
                                      C.c_void_p(base), ws.numel() - (base - ws.data_ptr()), stream),
                    "vmc_vit_forward",
                )
        return out

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: normalised fp32 pixel_values [F,3,224,224] (what the reference feeds ``visual_encoder``)
        -> fp32 [F, output_dim].  bf16 patchification happens here; callers that start from uint8
        should use ``ops.prologue`` + ``forward_patches`` and skip the fp32 image entirely."""
        if not x.is_cuda:
            raise _lib.VmcError("VisionTower.forward needs a CUDA tensor (no CPU fallback)")
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != self.input_resolution or x.shape[3] != self.input_resolution:
            raise ValueError(f"expected [F,3,{self.input_resolution},{self.input_resolution}] pixel_values")
        patches = ops.prologue(x.float(), wrap=False, dst="patch", patch=self.patch_size, normalised=True)
        return self.forward_patches(patches, x.shape[0])

    def flops_per_frame(self) -> float:
        """Algorithmic FLOPs per frame (2*MAC; SURVEY.md section 8d formula)."""
        L, d, h = self.tokens, self.width, self.heads
        per_layer = 2 * L * d * 3 * d + 2 * 2 * h * L * L * 64 + 2 * L * d * d + 2 * 2 * L * d * 4 * d
        patch = 2 * (L - 1) * 3 * self.patch_size**2 * d
        return patch + self.layers * per_layer + 2 * d * self.output_dim


def load_hf_vision_state_dict(tower: VisionTower, hf_sd: dict) -> None:
    """Load HF ``CLIPModel`` / ``CLIPVisionModelWithProjection`` weights (mapping of SURVEY.md App. A)."""
    d = tower.width
    pre = "vision_model."
    sd = {}
    sd["conv1.weight"] = hf_sd[pre + "embeddings.patch_embedding.weight"]
    sd["class_embedding"] = hf_sd[pre + "embeddings.class_embedding"]
    sd["positional_embedding"] = hf_sd[pre + "embeddings.position_embedding.weight"]
    sd["ln_pre.weight"], sd["ln_pre.bias"] = hf_sd[pre + "pre_layrnorm.weight"], hf_sd[pre + "pre_layrnorm.bias"]
    sd["ln_post.weight"], sd["ln_post.bias"] = hf_sd[pre + "post_layernorm.weight"], hf_sd[pre + "post_layernorm.bias"]
    sd["proj"] = hf_sd["visual_projection.weight"].t()
    for i in range(tower.layers):
        s, t = pre + f"encoder.layers.{i}.", f"transformer.resblocks.{i}."
        sd[t + "attn.in_proj_weight"] = torch.cat([hf_sd[s + f"self_attn.{n}_proj.weight"] for n in "qkv"], 0)
        sd[t + "attn.in_proj_bias"] = torch.cat([hf_sd[s + f"self_attn.{n}_proj.bias"] for n in "qkv"], 0)
        for a, b in [("attn.out_proj", "self_attn.out_proj"), ("ln_1", "layer_norm1"), ("ln_2", "layer_norm2"),
                     ("mlp.c_fc", "mlp.fc1"), ("mlp.c_proj", "mlp.fc2")]:
            sd[t + a + ".weight"], sd[t + a + ".bias"] = hf_sd[s + b + ".weight"], hf_sd[s + b + ".bias"]
    assert sd["conv1.weight"].shape[0] == d
    tower.load_state_dict(sd, strict=True)
