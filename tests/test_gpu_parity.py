"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle and the golden fixtures.

Bars (BASELINE.json north_star): bit-exact for the integer prologue / frame difference / indexing and
for the fp32 normalised tensor; per-frame embedding cosine >= 0.9995 and logit max-abs <= 1e-2 for
the bf16 stages.
"""
import math

import numpy as np
import pytest
import torch

import vimoclip_b200 as vmc
from oracle import clip_shim, losses as olosses, prologue, student as ostudent, tfam as otfam, weights
from vimoclip_b200 import ops

pytestmark = pytest.mark.gpu

COS_MIN = 0.9995  # north_star: per-frame embedding cosine
LOGIT_TOL = 1e-2  # north_star: logit max-abs


def _cos_min(a, b):
    return float(torch.nn.functional.cosine_similarity(a.double().cpu(), b.double().cpu(), dim=-1).min())


# ------------------------------------------------------------------------------------------------
# P1 prologue / frame difference: bit-exact
# ------------------------------------------------------------------------------------------------
def test_loaded_native_library(cuda_device):
    sm, major, minor = ops.device_info()
    assert major == 10, f"expected a Blackwell sm_100 device, got sm_{major}{minor}"
    assert sm > 0


def test_prologue_bit_exact_all_regimes(cuda_device, golden):
    g = golden("prologue.npz")
    u8 = torch.from_numpy(g["frame_u8"])[None].to(cuda_device)
    # regime A (uint8 input): integer wrap and fp32 normalise
    assert np.array_equal(ops.prologue(u8, wrap=True, dst="u8").cpu().numpy()[0], g["wrapA_u8"])
    assert np.array_equal(ops.prologue(u8, wrap=True, dst="f32").cpu().numpy()[0].view(np.uint32), g["normA"].view(np.uint32))
    # regime B (float [0,1]) and C (already normalised floats)
    fB = (u8.float() / 255.0).contiguous()
    assert np.array_equal(ops.prologue(fB, wrap=True, dst="u8").cpu().numpy(), u8.cpu().numpy())
    assert np.array_equal(ops.prologue(fB, wrap=True, dst="f32").cpu().numpy()[0].view(np.uint32), g["normB"].view(np.uint32))
    fC = torch.from_numpy(g["normB"])[None].to(cuda_device)
    assert np.array_equal(ops.prologue(fC, wrap=True, dst="u8").cpu().numpy()[0], g["wrapC_u8"])
    assert np.array_equal(ops.prologue(fC, wrap=True, dst="f32").cpu().numpy()[0].view(np.uint32), g["normC"].view(np.uint32))
    # no-wrap uint8 (HF processor path): (u8/255 - mean)/std
    ref = prologue.normalise_u8(g["frame_u8"][None])
    assert np.array_equal(ops.prologue(u8, wrap=False, dst="f32").cpu().numpy().view(np.uint32), ref.view(np.uint32))


@pytest.fixture(params=[0, 1, 2, 4], ids=["gather", "direct", "band", "gather16"])
def prologue_impl(request):
    ops.set_option(vmc._lib.OPT_PROLOGUE_IMPL, request.param)
    yield request.param
    ops.set_option(vmc._lib.OPT_PROLOGUE_IMPL, 0)


@pytest.mark.parametrize("patch", [32, 16, 14])
def test_prologue_patchify_bf16(cuda_device, prologue_impl, patch):
    gen = torch.Generator().manual_seed(5)
    u8 = torch.randint(0, 256, (3, 3, 224, 224), dtype=torch.uint8, generator=gen)
    got = ops.prologue(u8.to(cuda_device), wrap=True, dst="patch", patch=patch).cpu()
    ref32 = prologue.patchify(prologue.preprocess_frames(u8.numpy()), patch)
    ref = torch.from_numpy(ref32).to(torch.bfloat16)  # bf16 round-to-nearest-even of the bit-exact fp32 value
    k = 3 * patch * patch
    assert got.shape == (3 * (224 // patch) ** 2, ops.patch_ld(patch))
    assert torch.equal(got[:, :k].view(torch.int16), ref.view(torch.int16))
    assert torch.count_nonzero(got[:, k:]) == 0
    # no-wrap uint8 (HF processor path, extract_embeddings.py:89-93)
    gotN = ops.prologue(u8.to(cuda_device), wrap=False, dst="patch", patch=patch).cpu()
    refN = torch.from_numpy(prologue.patchify(prologue.normalise_u8(u8.numpy()), patch)).to(torch.bfloat16)
    assert torch.equal(gotN[:, :k].view(torch.int16), refN.view(torch.int16))
    assert torch.count_nonzero(gotN[:, k:]) == 0
    # float wrap regimes through the patch path
    fB = (u8.float() / 255.0).to(cuda_device)
    gotB = ops.prologue(fB, wrap=True, dst="patch", patch=patch).cpu()
    refB = torch.from_numpy(prologue.patchify(prologue.preprocess_frames(fB.cpu().numpy()), patch)).to(torch.bfloat16)
    assert torch.equal(gotB[:, :k].view(torch.int16), refB.view(torch.int16))


def test_frame_difference_bit_exact(cuda_device, golden, prologue_impl):
    g = golden("framediff.npz")
    frames = torch.from_numpy(g["frames"])[None].to(cuda_device)  # [1,5,48,64,3]
    diff, u8 = ops.frame_diff(frames, dst="u8")
    assert np.array_equal(diff.cpu().numpy()[0], g["diff"])
    want = prologue.to_pil_u8(np.repeat(g["diff"][:, None], 3, axis=1))  # 3 identical channels, regime A wrap
    assert np.array_equal(u8.cpu().numpy(), want)
    # random 224x224 clips, fp32 + patch outputs against the oracle
    rng = np.random.default_rng(11)
    bgr = rng.integers(0, 256, size=(2, 4, 224, 224, 3), dtype=np.uint8)
    d_ref = np.stack([prologue.frame_difference(c) for c in bgr])  # [2,3,224,224]
    diff, f32 = ops.frame_diff(torch.from_numpy(bgr).to(cuda_device), dst="f32")
    assert np.array_equal(diff.cpu().numpy(), d_ref)
    rep = np.repeat(d_ref.reshape(6, 1, 224, 224), 3, axis=1)
    n_ref = prologue.preprocess_frames(rep)
    assert np.array_equal(f32.cpu().numpy().view(np.uint32), n_ref.view(np.uint32))
    _, pt = ops.frame_diff(torch.from_numpy(bgr).to(cuda_device), dst="patch", patch=32, want_diff=False)
    assert torch.equal(pt.cpu().view(torch.int16), torch.from_numpy(prologue.patchify(n_ref, 32)).to(torch.bfloat16).view(torch.int16))


@pytest.mark.parametrize("H,W", [(360, 640), (240, 320), (500, 375), (224, 224), (300, 224), (1080, 1920), (120, 160), (100, 300),
                                 (223, 225), (64, 64), (30, 500)])
def test_resize_center_crop_bit_exact(cuda_device, golden, H, W):
    """Pillow bicubic Resize(224) + CenterCrop(224): bit-exact against the oracle (itself pinned to PIL), frames smaller
    than 224 included (scaled up: the short side becomes 224, so the crop never pads)."""
    from oracle import resize

    g = golden("resize.npz")
    tag = f"{H}x{W}"
    seed = int(g["seed_" + tag]) if "seed_" + tag in g.files else H * 3 + W
    img = np.random.default_rng(seed).integers(0, 256, size=(H, W, 3), dtype=np.uint8)
    chw = np.ascontiguousarray(img.transpose(2, 0, 1))[None]
    got = ops.resize_center_crop(torch.from_numpy(chw).to(cuda_device), wrap=False).cpu().numpy()[0]
    assert np.array_equal(got, resize.resize_center_crop_u8(chw[0]))
    if "pil_" + tag in g.files:
        assert np.array_equal(got, g["pil_" + tag])
    # student regimes: the wrap precedes the resize
    got_w = ops.resize_center_crop(torch.from_numpy(chw).to(cuda_device), wrap=True).cpu().numpy()[0]
    assert np.array_equal(got_w, resize.resize_center_crop_u8(prologue.to_pil_u8(chw[0])))
    f = (torch.from_numpy(chw).float() / 255.0).to(cuda_device)
    got_f = ops.resize_center_crop(f, wrap=True).cpu().numpy()[0]
    assert np.array_equal(got_f, got)  # regime B round-trips to the original uint8
    assert ops.resize_geometry(H, W)[:2] == resize.resized_size(H, W)


@pytest.mark.parametrize("H,W", [(360, 648), (227, 224), (360, 640), (500, 375)])
def test_resize_center_crop_hf_floor_offset(cuda_device, golden, H, W):
    """HF CLIPImageProcessor crop offset (margin // 2; extract_embeddings.py:18,91), against crops frozen from the installed
    processor where it differs from torchvision's round-half-to-even (margin 3 mod 4)."""
    from oracle import resize

    g = golden("resize.npz")
    tag = f"{H}x{W}"
    seed = int(g["hfseed_" + tag]) if "hfseed_" + tag in g.files else H * 5 + W
    img = np.random.default_rng(seed).integers(0, 256, size=(H, W, 3), dtype=np.uint8)
    chw = np.ascontiguousarray(img.transpose(2, 0, 1))[None]
    got = ops.resize_center_crop(torch.from_numpy(chw).to(cuda_device), wrap=False, hf_crop=True).cpu().numpy()[0]
    assert np.array_equal(got, resize.resize_center_crop_u8(chw[0], hf_crop=True))
    if "hf_" + tag in g.files:
        assert np.array_equal(got, g["hf_" + tag])
    feats = vmc.CLIPVisionFeatures("openai/clip-vit-base-patch32").to(cuda_device)
    a = feats.get_image_features_u8(torch.from_numpy(chw).to(cuda_device))
    b = feats.get_image_features_u8(torch.from_numpy(got[None]).to(cuda_device))
    assert torch.equal(a, b)  # the drop-in applies exactly this resize + crop


def test_prologue_full_size_properties(cuda_device):
    """BASELINE-size batch (256 clips x 16 frames): size-independent integer properties."""
    gen = torch.Generator(device="cuda").manual_seed(3)
    u8 = torch.randint(0, 256, (1024, 3, 224, 224), dtype=torch.uint8, device=cuda_device, generator=gen)
    w = ops.prologue(u8, wrap=True, dst="u8")
    assert torch.equal(ops.prologue(w, wrap=True, dst="u8"), u8)  # (-(-x)) mod 256 == x
    assert int(w.long().sum()) == int(((256 - u8.long()) % 256).sum())  # checksum of the closed form
    bgr = torch.randint(0, 256, (8, 17, 224, 224, 3), dtype=torch.uint8, device=cuda_device, generator=gen)
    same = bgr[:, :1].expand(-1, 17, -1, -1, -1).contiguous()
    d0, _ = ops.frame_diff(same)
    assert int(d0.max()) == 0  # identical frames -> zero difference
    d1, _ = ops.frame_diff(bgr)
    d2, _ = ops.frame_diff(bgr.flip(1).contiguous())
    assert torch.equal(d1, d2.flip(1))  # |a-b| is symmetric under time reversal


# ------------------------------------------------------------------------------------------------
# tcgen05 GEMM
# ------------------------------------------------------------------------------------------------
def _ref_gemm(a, w, bias, act, alpha, resid):
    y = a.float() @ w.float().t()
    if bias is not None:
        y = y + bias
    if act == ops.ACT_QUICKGELU:
        y = y * torch.sigmoid(1.702 * y)
    elif act == ops.ACT_GELU_ERF:
        y = torch.nn.functional.gelu(y)
    elif act == ops.ACT_RELU:
        y = torch.relu(y)
    y = y * alpha
    if resid is not None:
        y = y + resid
    return y


GEMM_CASES = [
    # M, N, K, act, bias, resid, out_dtype
    (128, 128, 64, ops.ACT_NONE, False, False, torch.float32),
    (100, 64, 128, ops.ACT_NONE, True, False, torch.float32),
    (333, 768, 768, ops.ACT_NONE, True, True, torch.float32),
    (1000, 2304, 768, ops.ACT_NONE, True, False, torch.bfloat16),
    (1576, 3072, 768, ops.ACT_QUICKGELU, True, False, torch.bfloat16),
    (1576, 768, 3072, ops.ACT_NONE, True, True, torch.float32),
    (40000, 768, 768, ops.ACT_NONE, True, False, torch.bfloat16),  # 128x256 tiles, several tiles per CTA
    (512, 140, 256, ops.ACT_NONE, True, False, torch.float32),  # ragged N
    (512, 512, 512, ops.ACT_GELU_ERF, True, False, torch.bfloat16),
    (64, 256, 512, ops.ACT_RELU, True, False, torch.bfloat16),
    (129, 768, 64, ops.ACT_NONE, True, True, torch.float32),  # second CTA of the pair owns a single row
    (50000, 2304, 768, ops.ACT_NONE, True, False, torch.bfloat16),  # many pair tiles per cluster
    (1000, 36, 128, ops.ACT_NONE, True, True, torch.float32),  # ragged 32-column chunk
]


@pytest.fixture(params=[2, 1], ids=["pair", "single"])
def gemm_impl(request):
    ops.set_option(vmc._lib.OPT_GEMM_IMPL, request.param)
    yield request.param
    ops.set_option(vmc._lib.OPT_GEMM_IMPL, 0)


@pytest.mark.parametrize("M,N,K,act,use_bias,use_resid,odt", GEMM_CASES)
def test_gemm_against_fp32(cuda_device, gemm_impl, M, N, K, act, use_bias, use_resid, odt):
    gen = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    a = (torch.randn(M, K, device=cuda_device, generator=gen)).to(torch.bfloat16)
    w = (torch.randn(N, K, device=cuda_device, generator=gen) * K**-0.5).to(torch.bfloat16)
    bias = torch.randn(N, device=cuda_device, generator=gen) if use_bias else None
    resid = torch.randn(M, N, device=cuda_device, generator=gen) if use_resid else None
    alpha = 0.5 if use_resid else 1.0
    got = ops.gemm(a, w, bias=bias, act=act, alpha=alpha, resid=resid, out_dtype=odt).float()
    ref = _ref_gemm(a, w, bias, act, alpha, resid)
    tol = 2e-2 if odt == torch.bfloat16 else 2e-4
    err = (got - ref).abs().max().item()
    assert err <= tol * max(1.0, ref.abs().max().item()), f"max abs err {err}"


def test_gemm_patch_embed_rowgroup_and_padded_k(cuda_device, gemm_impl):
    """Patch-embed epilogue: token rows scattered past each frame's CLS row, pos-emb added; K = 588 (ViT-L/14)."""
    gen = torch.Generator(device="cuda").manual_seed(1)
    F_, n, d, K, ld = 3, 256, 1024, 588, 592
    a = torch.zeros(F_ * n, ld, device=cuda_device, dtype=torch.bfloat16)
    a[:, :K] = torch.randn(F_ * n, K, device=cuda_device, generator=gen).to(torch.bfloat16)
    a[:, K:] = 7.0  # pad columns must be ignored: K < ld
    w = torch.zeros(d, ld, device=cuda_device, dtype=torch.bfloat16)
    w[:, :K] = (torch.randn(d, K, device=cuda_device, generator=gen) * K**-0.5).to(torch.bfloat16)
    w[:, K:] = 3.0
    pos = torch.randn(n + 1, d, device=cuda_device, generator=gen)
    x = torch.full((F_ * (n + 1), d), -5.0, device=cuda_device)
    ops.gemm(a, w, resid=pos, out=x, k=K, row_group=n)
    ref = (a[:, :K].float() @ w[:, :K].float().t()).view(F_, n, d) + pos[1:][None]
    xv = x.view(F_, n + 1, d)
    assert torch.all(xv[:, 0] == -5.0)  # CLS rows untouched
    assert (xv[:, 1:] - ref).abs().max().item() < 2e-4 * max(1.0, ref.abs().max().item())


def test_vit_tower_layernorm_variants(cuda_device):
    """Tower variants (vmc_vit_model.ln_mode): 4 = fp32 residual stream with separate LayerNorm kernels is the yardstick;
    5 / 3 fold the LayerNorms into the consuming GEMMs (different rounding points, same mathematics); 6, the default, also
    keeps the residual stream in bf16.  Per-model selectors: no process-wide state is touched."""
    torch.manual_seed(0)
    tower = vmc.VisionTower.from_name("ViT-B/16").to(cuda_device)
    gen = torch.Generator().manual_seed(6)
    u8 = torch.randint(0, 256, (80, 3, 224, 224), dtype=torch.uint8, generator=gen).to(cuda_device)  # 80*197 rows: 256-wide tiles
    patches = ops.prologue(u8, wrap=False, dst="patch", patch=16)
    tower.last_block_cls = 2
    tower.ln_mode = 4
    separate = tower.forward_patches(patches, 80).clone()
    for mode, bar in ((5, 0.99998), (3, 0.99998), (6, 0.9999), (0, 0.9999)):
        tower.ln_mode = mode
        got = tower.forward_patches(patches, 80)
        cos = torch.nn.functional.cosine_similarity(got.double(), separate.double(), dim=-1).min().item()
        assert cos >= bar, (mode, cos)
    tower.ln_mode = 0
    a = tower.forward_patches(patches, 80).clone()
    tower.ln_mode = 6
    assert torch.equal(a, tower.forward_patches(patches, 80))  # 0 = library default = 6


@pytest.mark.parametrize("M,N,K", [(394, 768, 768), (50000, 768, 3072), (40000, 1024, 1024), (700, 512, 2048)])
def test_gemm_bf16_residual_stream(cuda_device, M, N, K):
    """The residual GEMM of the default tower: out = bf16(acc + bias + bf16 resid), in place, plus the per-row partial
    (sum, sum of squares) of the ROUNDED values; a LayerNorm-folded consumer GEMM on those rows matches fp32 LayerNorm."""
    gen = torch.Generator(device="cuda").manual_seed(M + K)
    a0 = torch.randn(M, K, device=cuda_device, generator=gen).to(torch.bfloat16)
    w0 = (torch.randn(N, K, device=cuda_device, generator=gen) * K**-0.5).to(torch.bfloat16)
    b0 = torch.randn(N, device=cuda_device, generator=gen)
    x = (torch.randn(M, N, device=cuda_device, generator=gen) * 2 + 0.5)
    x[:, 7] += 30.0
    x16 = x.to(torch.bfloat16)
    parts = ops.gemm_stats_parts(M, N)
    stats = torch.zeros(parts, M, 2, device=cuda_device)
    xs = x16.clone()
    ops.gemm(a0, w0, bias=b0, resid=xs, out=xs, emit_stats=(None, stats))
    ref = (a0.float() @ w0.float().t() + b0 + x16.float())
    # bf16 rounding of the reference: allow one ulp of disagreement where the fp32 sums differ in the last bits
    assert (xs.float() - ref).abs().max().item() <= 2.0 ** -7 * ref.abs().max().item()
    assert (xs.float() - ref.to(torch.bfloat16).float()).abs().mean().item() < 1e-3
    st = stats.sum(0)
    xr = xs.float()
    assert torch.allclose(st[:, 0], xr.sum(1), rtol=1e-5, atol=1e-2)
    assert torch.allclose(st[:, 1], (xr * xr).sum(1), rtol=1e-5, atol=1e-2)
    # consumer: LayerNorm folded into a GEMM that reads the very same buffer
    Nc = 256
    gamma = torch.randn(N, device=cuda_device, generator=gen) * 0.3 + 1.0
    beta = torch.randn(N, device=cuda_device, generator=gen) * 0.2
    w1 = torch.randn(Nc, N, device=cuda_device, generator=gen) * N**-0.5
    b1 = torch.randn(Nc, device=cuda_device, generator=gen)
    wf = (w1 * gamma[None, :]).to(torch.bfloat16)
    got = ops.gemm(xs, wf, bias=b1 + w1 @ beta, fold=(stats, wf.float().sum(1), 1e-5)).float()
    want = torch.nn.functional.layer_norm(xr, (N,), gamma, beta, 1e-5) @ w1.t() + b1
    assert (got - want).abs().max().item() < 3e-2 * max(1.0, want.abs().max().item())
    # fp32 output with a bf16 residual (the CLS rows of the last block): dtype switch of the residual read only
    y = ops.gemm(a0, w0, bias=b0, resid=x16, out_dtype=torch.float32)
    assert (y - ref).abs().max().item() < 2e-4 * max(1.0, ref.abs().max().item())


def test_gemm_linearity_full_size(cuda_device, gemm_impl):
    """BASELINE-size GEMM (64 frames x 197 tokens, fc1): G(a1 + a2) == G(a1) + G(a2) on exactly representable inputs."""
    gen = torch.Generator(device="cuda").manual_seed(9)
    M, N, K = 64 * 197, 3072, 768
    a1 = torch.randint(-4, 5, (M, K), device=cuda_device, generator=gen).to(torch.bfloat16)
    a2 = torch.randint(-4, 5, (M, K), device=cuda_device, generator=gen).to(torch.bfloat16)
    w = torch.randint(-2, 3, (N, K), device=cuda_device, generator=gen).to(torch.bfloat16)
    y1 = ops.gemm(a1, w, out_dtype=torch.float32)
    y2 = ops.gemm(a2, w, out_dtype=torch.float32)
    y12 = ops.gemm((a1 + a2), w, out_dtype=torch.float32)
    assert torch.equal(y12, y1 + y2)  # small integers: every product and sum is exact in fp32
    assert torch.equal(y1, a1.float() @ w.float().t())


# ------------------------------------------------------------------------------------------------
# LayerNorm / attention / small kernels
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d", [512, 768, 1024])
def test_layernorm(cuda_device, d):
    gen = torch.Generator(device="cuda").manual_seed(d)
    x = torch.randn(777, d, device=cuda_device, generator=gen) * 3 + 1
    g_ = torch.randn(d, device=cuda_device, generator=gen)
    b_ = torch.randn(d, device=cuda_device, generator=gen)
    y32, y16 = ops.layernorm(x, g_, b_, want32=True, want16=True)
    ref = torch.nn.functional.layer_norm(x, (d,), g_, b_, 1e-5)
    assert (y32 - ref).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item())
    assert torch.equal(y16, y32.to(torch.bfloat16))
    # bf16 input rows (the tower's bf16 residual stream) and the statistics of the rounded output
    xb = x.to(torch.bfloat16)
    stats = torch.zeros(777, 2, device=cuda_device)
    z32, z16 = ops.layernorm(xb, g_, b_, want32=True, want16=True, stats=stats, stats_rounded=True)
    refb = torch.nn.functional.layer_norm(xb.float(), (d,), g_, b_, 1e-5)
    assert (z32 - refb).abs().max().item() < 2e-5 * max(1.0, refb.abs().max().item())
    assert torch.equal(z16, z32.to(torch.bfloat16))
    zr = z16.float()
    assert torch.allclose(stats[:, 0], zr.sum(1), rtol=1e-5, atol=1e-3)
    assert torch.allclose(stats[:, 1], (zr * zr).sum(1), rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("M,N,K,col0,pad", [(1000, 768, 256, 0, 0), (1000, 768, 256, 64, 200), (333, 2304, 128, 8, 8), (4097, 320, 192, 0, 8),
                                            (70, 96, 64, 32, 40)])
def test_gemm_tma_epilogues_strided_views(cuda_device, M, N, K, col0, pad):
    """The bf16 epilogues move their rows by TMA (store; load + store for the bf16 residual stream): outputs and residuals that
    are column slices of wider buffers (row stride > N, base offset), ragged last row tiles, out-of-place residual GEMMs with
    different row strides, the untouched columns around the slice, and the LDS + STG fallback for strides TMA cannot take."""
    gen = torch.Generator(device="cuda").manual_seed(M + N + col0)
    a = torch.randn(M, K, device=cuda_device, generator=gen).to(torch.bfloat16)
    w = (torch.randn(N, K, device=cuda_device, generator=gen) * K**-0.5).to(torch.bfloat16)
    b = torch.randn(N, device=cuda_device, generator=gen)
    ref = a.float() @ w.float().t() + b
    tol = 2e-2 * max(1.0, ref.abs().max().item())
    for act in (ops.ACT_NONE, ops.ACT_QUICKGELU):
        wide = torch.full((M, col0 + N + pad), 7.0, device=cuda_device, dtype=torch.bfloat16)
        out = wide[:, col0:col0 + N]
        ops.gemm(a, w, bias=b, act=act, out=out)
        want = ref if act == ops.ACT_NONE else ref * torch.sigmoid(1.702 * ref)
        assert (out.float() - want).abs().max().item() <= tol
        assert torch.all(wide[:, :col0] == 7.0) and torch.all(wide[:, col0 + N:] == 7.0)  # nothing written outside the slice
        ops.set_option(vmc._lib.OPT_GEMM_IMPL, 3)  # LDS + STG form of the same epilogue: same numbers
        try:
            chk = ops.gemm(a, w, bias=b, act=act)
        finally:
            ops.set_option(vmc._lib.OPT_GEMM_IMPL, 0)
        assert torch.equal(chk, out.contiguous())
    with pytest.raises(vmc._lib.VmcError):  # the C-ABI contract: 16-byte aligned output rows
        ops.gemm(a, w, bias=b, out=torch.empty(M, N + 4, device=cuda_device, dtype=torch.bfloat16)[:, :N])
    if N % 128 == 0:
        # bf16 residual stream, out of place, residual and output with different row strides and offsets
        x16 = (torch.randn(M, N + 24, device=cuda_device, generator=gen) * 2).to(torch.bfloat16)
        resid = x16[:, 16:16 + N]
        wide = torch.full((M, col0 + N + pad), 7.0, device=cuda_device, dtype=torch.bfloat16)
        out = wide[:, col0:col0 + N]
        stats = torch.zeros(ops.gemm_stats_parts(M, N), M, 2, device=cuda_device)
        ops.gemm(a, w, bias=b, resid=resid, out=out, emit_stats=(None, stats))
        want = ref + resid.float()
        assert (out.float() - want).abs().max().item() <= 2.0 ** -7 * max(1.0, want.abs().max().item())
        assert torch.all(wide[:, :col0] == 7.0) and torch.all(wide[:, col0 + N:] == 7.0)
        st, xr = stats.sum(0), out.float()
        assert torch.allclose(st[:, 0], xr.sum(1), rtol=1e-5, atol=1e-2)
        assert torch.allclose(st[:, 1], (xr * xr).sum(1), rtol=1e-5, atol=1e-2)


def test_gemm_tma_epilogues_deterministic_at_scale(cuda_device):
    """Many tiles per CTA pair (the staging tiles and residual boxes are reused every 32-column chunk): the TMA epilogues give
    the same bits run to run, the same bits as the LDS + STG form, and the in-place residual GEMM the same as out of place."""
    gen = torch.Generator(device="cuda").manual_seed(5)
    M, d = 150001, 768
    x = torch.randn(M, d, device=cuda_device, generator=gen).to(torch.bfloat16)
    stats = torch.zeros(ops.gemm_stats_parts(M, d), M, 2, device=cuda_device)
    xf = x.float()
    st_in = torch.stack([xf.sum(1), (xf * xf).sum(1)], 1)[None].contiguous()
    del xf
    for N, act in ((2304, ops.ACT_NONE), (3072, ops.ACT_QUICKGELU)):
        w = (torch.randn(N, d, device=cuda_device, generator=gen) * d**-0.5).to(torch.bfloat16)
        b = torch.randn(N, device=cuda_device, generator=gen)
        cs = w.float().sum(1)
        first = ops.gemm(x, w, bias=b, act=act, fold=(st_in, cs, 1e-5))
        for _ in range(2):
            assert torch.equal(first, ops.gemm(x, w, bias=b, act=act, fold=(st_in, cs, 1e-5)))
        ops.set_option(vmc._lib.OPT_GEMM_IMPL, 3)
        try:
            assert torch.equal(first, ops.gemm(x, w, bias=b, act=act, fold=(st_in, cs, 1e-5)))
        finally:
            ops.set_option(vmc._lib.OPT_GEMM_IMPL, 0)
        del first
    for K in (768, 3072):
        a = torch.randn(M, K, device=cuda_device, generator=gen).to(torch.bfloat16)
        w = (torch.randn(d, K, device=cuda_device, generator=gen) * K**-0.5).to(torch.bfloat16)
        b = torch.randn(d, device=cuda_device, generator=gen)
        out = torch.empty_like(x)
        ops.gemm(a, w, bias=b, resid=x, out=out, emit_stats=(None, stats))
        st0 = stats.clone()
        for _ in range(2):
            xs = x.clone()
            ops.gemm(a, w, bias=b, resid=xs, out=xs, emit_stats=(None, stats))  # in place
            assert torch.equal(xs, out) and torch.equal(stats, st0)
        del a, out


@pytest.mark.parametrize("M,d,N,act", [(394, 768, 2304, 0), (50000, 768, 3072, 1), (40000, 1024, 3072, 0), (700, 512, 1536, 1)])
def test_gemm_layernorm_fold(cuda_device, M, d, N, act):
    """LayerNorm folded into the consuming GEMM: a producer GEMM (bias + fp32 residual) emits bf16 rows + partial row
    statistics, the consumer runs on the raw rows with gamma-folded weights (include/vimoclip_b200.h).  Checked against
    fp32 LayerNorm + matmul on the same bf16 operands."""
    gen = torch.Generator(device="cuda").manual_seed(M + N)
    a0 = (torch.randn(M, d, device=cuda_device, generator=gen)).to(torch.bfloat16)
    w0 = (torch.randn(d, d, device=cuda_device, generator=gen) * d**-0.5).to(torch.bfloat16)
    b0 = torch.randn(d, device=cuda_device, generator=gen)
    x = torch.randn(M, d, device=cuda_device, generator=gen) * 2 + 0.5
    x[:, 7] += 30.0  # an outlier channel and a non-zero mean, as in real residual streams
    parts = ops.gemm_stats_parts(M, d)
    raw16 = torch.empty(M, d, dtype=torch.bfloat16, device=cuda_device)
    stats = torch.zeros(parts, M, 2, device=cuda_device)
    xnew = x.clone()
    ops.gemm(a0, w0, bias=b0, resid=xnew, out=xnew, emit_stats=(raw16, stats))
    ref_x = a0.float() @ w0.float().t() + b0 + x
    assert (xnew - ref_x).abs().max().item() < 2e-3
    assert torch.equal(raw16, xnew.to(torch.bfloat16))
    st = stats.sum(0)
    assert torch.allclose(st[:, 0], xnew.sum(1), rtol=1e-5, atol=1e-2)
    assert torch.allclose(st[:, 1], (xnew * xnew).sum(1), rtol=1e-5, atol=1e-2)
    # consumer
    gamma = torch.randn(d, device=cuda_device, generator=gen) * 0.3 + 1.0
    beta = torch.randn(d, device=cuda_device, generator=gen) * 0.2
    w1 = torch.randn(N, d, device=cuda_device, generator=gen) * d**-0.5
    b1 = torch.randn(N, device=cuda_device, generator=gen)
    wf = (w1 * gamma[None, :]).to(torch.bfloat16)
    colsum = wf.float().sum(1)
    bf = b1 + w1 @ beta
    got = ops.gemm(raw16, wf, bias=bf, act=act, fold=(stats, colsum, 1e-5)).float()
    y = torch.nn.functional.layer_norm(xnew, (d,), gamma, beta, 1e-5)
    ref = y @ w1.t() + b1
    if act:
        ref = ref * torch.sigmoid(1.702 * ref)
    # reference path of the library: LayerNorm kernel -> bf16 -> plain GEMM
    _, y16 = ops.layernorm(xnew, gamma, beta)
    base = ops.gemm(y16, w1.to(torch.bfloat16), bias=b1, act=act).float()
    err_fold = (got - ref).abs().max().item()
    err_base = (base - ref).abs().max().item()
    assert err_fold < max(2.0 * err_base, 3e-2), (err_fold, err_base)


@pytest.mark.parametrize("impl", [8, 7, 6, 5, 2])
@pytest.mark.parametrize("L,heads,F_", [(50, 12, 5), (197, 12, 3), (257, 16, 2), (128, 2, 2), (16, 1, 1), (129, 3, 2), (272, 1, 1), (197, 12, 40), (256, 4, 75), (200, 1, 1),
                                        (257, 16, 60), (226, 2, 3), (241, 3, 7), (145, 1, 2), (50, 12, 200), (50, 1, 7), (64, 3, 5), (33, 2, 3), (7, 1, 1)])
def test_attention_vit(cuda_device, L, heads, F_, impl):
    gen = torch.Generator(device="cuda").manual_seed(L)
    d = heads * 64
    qkv = (torch.randn(F_ * L, 3 * d, device=cuda_device, generator=gen) * 1.5).to(torch.bfloat16)
    got = ops.attention_vit(qkv, F_, L, heads, impl=impl).float().view(F_, L, heads, 64)
    q, k, v = qkv.float().view(F_, L, 3, heads, 64).unbind(2)
    s = torch.einsum("flhd,fmhd->fhlm", q, k) / 8.0
    ref = torch.einsum("fhlm,fmhd->flhd", torch.softmax(s, -1), v)
    err = (got - ref).abs().max().item()
    assert err < 2e-2, f"max abs err {err}"


@pytest.mark.parametrize("L,heads,F_", [(197, 12, 9), (50, 12, 33), (257, 16, 5), (1, 2, 3), (17, 1, 2), (8192, 1, 1)])
def test_attention_cls_row(cuda_device, L, heads, F_):
    """The last block's attention on the CLS query only (vmc_attention_cls) against fp32 softmax attention of that row."""
    gen = torch.Generator(device="cuda").manual_seed(L + heads)
    d = heads * 64
    q = (torch.randn(F_, d, device=cuda_device, generator=gen) * 2).to(torch.bfloat16)
    kv = (torch.randn(F_ * L, 2 * d, device=cuda_device, generator=gen) * 1.5).to(torch.bfloat16)
    got = ops.attention_cls(q, kv, F_, L, heads).float().view(F_, heads, 64)
    k, v = kv.float().view(F_, L, 2, heads, 64).unbind(2)
    s = torch.einsum("fhd,fmhd->fhm", q.float().view(F_, heads, 64), k) / 8.0
    ref = torch.einsum("fhm,fmhd->fhd", torch.softmax(s, -1), v)
    assert (got - ref).abs().max().item() < 2e-2


@pytest.mark.parametrize("impl,L", [(5, 197), (5, 224), (5, 160), (2, 197), (6, 257), (6, 197), (7, 50), (8, 50)])
def test_attention_vit_peaky_and_shifted_scores(cuda_device, impl, L):
    """Large score spread and a large common offset: the single-pass softmax (stabiliser = max of the
    first 32 keys) must stay as accurate as the exact-max reference.  (impl 6: the CLS token is handled outside the
    tensor-core pipeline -- as a key with the row's stabiliser, as a query in fp32.)"""
    gen = torch.Generator(device="cuda").manual_seed(77)
    F_, heads = 3, 2
    d = heads * 64
    qkv = torch.randn(F_ * L, 3 * d, device=cuda_device, generator=gen)
    qkv[:, :d] *= 6.0  # peaky rows: scaled scores with std ~ 6, max - first-chunk max up to ~15
    qkv[::7, d:2 * d] += 1.0  # some keys systematically preferred
    qkv = qkv.to(torch.bfloat16)
    got = ops.attention_vit(qkv, F_, L, heads, impl=impl).float().view(F_, L, heads, 64)
    q, k, v = qkv.float().view(F_, L, 3, heads, 64).unbind(2)
    s = torch.einsum("flhd,fmhd->fhlm", q, k) / 8.0
    ref = torch.einsum("fhlm,fmhd->flhd", torch.softmax(s, -1), v)
    assert torch.isfinite(got).all()
    err = (got - ref).abs().max().item()
    assert err < 3e-2, f"max abs err {err}"


@pytest.mark.parametrize("impl,L,heads,F_", [(7, 50, 12, 700), (8, 50, 12, 700), (5, 197, 12, 300), (5, 160, 12, 200), (5, 224, 3, 500), (5, 129, 2, 400), (5, 208, 5, 333), (6, 257, 16, 260), (6, 230, 4, 500)])
def test_attention_vit_persistent_kernels_deterministic_at_scale(cuda_device, impl, L, heads, F_):
    """Many items per CTA (ring slots, TMEM tiles and hand-over buffers reused dozens of times): the persistent kernels
    must give bit-identical results run to run (a race shows up as run-to-run noise) and agree with the first-generation
    per-item kernel (impl 2) on the same input."""
    gen = torch.Generator(device="cuda").manual_seed(impl * 1000 + L)
    d = heads * 64
    qkv = (torch.randn(F_ * L, 3 * d, device=cuda_device, generator=gen) * 1.2).to(torch.bfloat16)
    first = ops.attention_vit(qkv, F_, L, heads, impl=impl)
    for _ in range(3):
        again = ops.attention_vit(qkv, F_, L, heads, impl=impl)
        assert torch.equal(first.view(torch.int16), again.view(torch.int16))
    base = ops.attention_vit(qkv, F_, L, heads, impl=2)
    # two bf16-rounded outputs: allow one bf16 ulp (2^-8 relative) of the value on top of the absolute bar
    assert bool(((first.float() - base.float()).abs() <= 1e-2 + 2.0 ** -7 * base.float().abs()).all())


@pytest.mark.parametrize("B,Tq,Tk", [(2, 16, 15), (3, 40, 100), (1, 5, 70)])
def test_attention_masked(cuda_device, B, Tq, Tk):
    gen = torch.Generator(device="cuda").manual_seed(Tk)
    h, d = 8, 512
    q = torch.randn(B * Tq, d, device=cuda_device, generator=gen)
    kv = torch.randn(B * Tk, 2 * d, device=cuda_device, generator=gen)
    valid = torch.rand(B, Tk, device=cuda_device, generator=gen) > 0.3
    valid[:, 0] = True
    got = ops.attention_masked(q, kv[:, :d], kv[:, d:], valid, B, Tq, Tk, h).float()
    qh = q.view(B, Tq, h, 64).transpose(1, 2)
    kh = kv[:, :d].reshape(B, Tk, h, 64).transpose(1, 2)
    vh = kv[:, d:].reshape(B, Tk, h, 64).transpose(1, 2)
    s = (qh @ kh.transpose(-1, -2)) / 8.0
    s = s.masked_fill(~valid[:, None, None, :], float("-inf"))
    ref = (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(B * Tq, d)
    assert (got - ref).abs().max().item() < 1e-2  # bf16 output rounding
    got2 = ops.attention_masked(q, kv[:, :d], kv[:, d:], None, B, Tq, Tk, h).float()
    ref2 = (torch.softmax((qh @ kh.transpose(-1, -2)) / 8.0, -1) @ vh).transpose(1, 2).reshape(B * Tq, d)
    assert (got2 - ref2).abs().max().item() < 1e-2


def test_small_kernels(cuda_device, golden):
    gen = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(6, 16, 512, device=cuda_device, generator=gen)
    m32, m16 = ops.mean_rows(x, want32=True, want16=True)
    assert (m32 - x.mean(1)).abs().max().item() < 1e-6
    assert torch.equal(m16, m32.to(torch.bfloat16))
    assert torch.equal(ops.cast_bf16(x.view(-1, 512)), x.view(-1, 512).to(torch.bfloat16))
    g = golden("losses.npz")
    s, t, t2 = (torch.from_numpy(g[k]).to(cuda_device) for k in ("s", "t", "t2"))
    assert abs(float(vmc.distillation_loss(s, t, "cosine")) - float(g["cos"])) < 1e-5
    assert abs(float(vmc.distillation_loss(s, t2, "cosine")) - float(g["cos2"])) < 1e-5
    assert abs(float(olosses.distillation_loss(s.cpu(), t.cpu(), "cosine")) - float(g["cos"])) < 1e-6


# ------------------------------------------------------------------------------------------------
# whole towers / drop-in modules against the oracle and the golden fixtures
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,nframes", [("ViT-B/32", 3), ("ViT-B/16", 2), ("ViT-L/14", 1)])
def test_vit_tower_against_oracle_and_hf_golden(cuda_device, golden, name, nframes):
    g = golden("vit_hf.npz")
    tag = name.replace("/", "").replace("-", "").lower()
    oracle = clip_shim.build_visual(name, seed=0)
    feats = vmc.CLIPVisionFeatures(name)
    feats.visual.load_state_dict(oracle.state_dict(), strict=True)
    feats = feats.to(cuda_device)
    gen = torch.Generator().manual_seed(4321)
    u8 = torch.randint(0, 256, (nframes, 3, 224, 224), dtype=torch.uint8, generator=gen)
    got = feats.get_image_features_u8(u8.to(cuda_device))
    ref = torch.from_numpy(g[tag + "_hf"])  # HF CLIPVisionModelWithProjection output (fp32, CPU)
    assert _cos_min(got, ref) >= COS_MIN, _cos_min(got, ref)
    x = torch.from_numpy(prologue.normalise_u8(u8.numpy()))
    with torch.no_grad():
        ref2 = oracle(x)
    assert _cos_min(got, ref2) >= COS_MIN
    # the reference-facing call: already-normalised fp32 pixel_values
    got2 = feats.get_image_features(x.to(cuda_device))
    assert _cos_min(got2, ref) >= COS_MIN
    assert (got.cpu() - ref).abs().max().item() < 0.05 * ref.abs().max().item()


def test_student_config1_against_reference_golden(cuda_device, golden):
    g = golden("student.npz")
    oracle = ostudent.StudentOracle("ViT-B/32", seed=0)
    weights.randomise_heads_(oracle, 0)
    ours = vmc.FrameDiffStudentModel("ViT-B/32", device=cuda_device, num_classes=140, alpha=0.1)
    ours.load_state_dict(oracle.state_dict(), strict=True)
    gen = torch.Generator().manual_seed(1234)
    frames = torch.randint(0, 256, (2, 16, 3, 224, 224), dtype=torch.uint8, generator=gen)
    emb, dis, logits = ours(frames)  # CPU uint8 input, like the reference's loaders hand over
    assert emb.shape == (2, 16, 512) and dis.shape == (2, 16, 512) and logits.shape == (2, 140)
    assert _cos_min(emb, torch.from_numpy(g["fd_b32_emb"])) >= COS_MIN
    assert _cos_min(dis, torch.from_numpy(g["fd_b32_distill"])) >= COS_MIN
    assert (logits.cpu() - torch.from_numpy(g["fd_b32_logits"])).abs().max().item() <= LOGIT_TOL
    # float regimes B and C
    gen = torch.Generator().manual_seed(99)
    u8 = torch.randint(0, 256, (1, 2, 3, 224, 224), dtype=torch.uint8, generator=gen)
    embB, _, logB = ours(u8.float() / 255.0)
    assert _cos_min(embB, torch.from_numpy(g["regB_emb"])) >= COS_MIN
    assert (logB.cpu() - torch.from_numpy(g["regB_logits"])).abs().max().item() <= LOGIT_TOL
    mean = torch.tensor(clip_shim.CLIP_MEAN).view(1, 1, 3, 1, 1)
    std = torch.tensor(clip_shim.CLIP_STD).view(1, 1, 3, 1, 1)
    embC, _, logC = ours((u8.float() / 255.0 - mean) / std)
    assert _cos_min(embC, torch.from_numpy(g["regC_emb"])) >= COS_MIN
    assert (logC.cpu() - torch.from_numpy(g["regC_logits"])).abs().max().item() <= LOGIT_TOL
    # 640x360 frames: wrap -> Pillow bicubic resize -> centre crop -> normalise, against the reference file's output
    gr = golden("resize.npz")
    gen = torch.Generator().manual_seed(31)
    big = torch.randint(0, 256, (1, 2, 3, 360, 640), dtype=torch.uint8, generator=gen)
    embR, _, logR = ours(big)
    assert _cos_min(embR, torch.from_numpy(gr["student_emb"])) >= COS_MIN
    assert (logR.cpu() - torch.from_numpy(gr["student_logits"])).abs().max().item() <= LOGIT_TOL
    # frames smaller than 224 are scaled up by the same resampler (the oracle applies PIL's arithmetic)
    small = torch.randint(0, 256, (1, 2, 3, 120, 160), dtype=torch.uint8, generator=gen)
    embS, _, logS = ours(small)
    with torch.no_grad():
        eS, _, lS = oracle(small)
    assert _cos_min(embS, eS) >= COS_MIN and (logS.cpu() - lS).abs().max().item() <= LOGIT_TOL


MODES = {
    "cross": dict(),
    "cross_pe": dict(use_pe=True),
    "rgb_only": dict(use_only_rgb=True),
    "flow_only": dict(use_only_flow=True),
    "concat_t": dict(use_cross_attention=False, concat_dim=1),
    "concat_e": dict(use_cross_attention=False, concat_dim=-1),
}


@pytest.mark.parametrize("fused", [True, False], ids=["fused", "batched"])
@pytest.mark.parametrize("tag", list(MODES))
def test_tfam_config1_against_reference_golden(cuda_device, golden, tag, fused):
    """Config 1 (B = 2, ragged masks), all six fusion modes, against logits of the reference file itself.  fused: the ONE
    kernel per forward (default); batched: the split-bf16 tcgen05 GEMM path (other geometries / long clips)."""
    g = golden("tfam.npz")
    oracle = otfam.TfamOracle(**MODES[tag]).eval()
    weights.randomise_tfam_(oracle, 0)
    ours = vmc.AMO_CLIP(device=cuda_device, **MODES[tag])
    ours.load_state_dict(oracle.state_dict(), strict=True)
    ours = ours.to(cuda_device).eval()
    ours.fused = fused
    rgb, mot = torch.from_numpy(g["rgb"]).to(cuda_device), torch.from_numpy(g["motion"]).to(cuda_device)
    mr, mm = torch.from_numpy(g["mask_rgb"]).to(cuda_device), torch.from_numpy(g["mask_mot"]).to(cuda_device)
    logits = ours(rgb.clone(), mot.clone(), mr, mm)  # warm: packs the weights
    ops.reset_launch_count()
    logits = ours(rgb.clone(), mot.clone(), mr, mm)
    launches = ops.launch_count()
    err = (logits.cpu() - torch.from_numpy(g[tag + "_logits"])).abs().max().item()
    assert logits.shape == (2, 140) and err <= LOGIT_TOL, f"{tag}: logit max-abs {err}"
    if fused:  # north star piece 4: one fused kernel (+ the projection GEMM and its cast in the embedding-concat mode)
        assert launches == (3 if tag == "concat_e" else 1), launches


@pytest.mark.parametrize("B,T,Tm", [(256, 16, 15), (33, 16, 16), (5, 32, 31), (7, 9, 4), (40, 1, 1), (3, 20, 32), (70, 12, 16)])
def test_tfam_fused_kernel_shapes_against_oracle(cuda_device, B, T, Tm):
    """The fused kernel against the fp32 oracle over batch sizes (one clip per cluster / two clips per pass, odd tails),
    sequence lengths (one and two M tiles) and ragged masks; equal to the batched path within the same bar; deterministic."""
    oracle = otfam.TfamOracle().eval()
    weights.randomise_tfam_(oracle, 5)
    ours = vmc.AMO_CLIP(device=cuda_device)
    ours.load_state_dict(oracle.state_dict(), strict=True)
    ours = ours.to(cuda_device).eval()
    gen = torch.Generator().manual_seed(B * 100 + T)
    rgb, mot = torch.randn(B, T, 512, generator=gen), torch.randn(B, Tm, 512, generator=gen)
    lr = torch.randint(1, T + 1, (B,), generator=gen)
    lm = torch.randint(1, Tm + 1, (B,), generator=gen)
    mr = torch.arange(T)[None] < lr[:, None]
    mm = torch.arange(Tm)[None] < lm[:, None]
    ref = oracle(rgb, mot, mr, mm)
    args = (rgb.to(cuda_device), mot.to(cuda_device), mr.to(cuda_device), mm.to(cuda_device))
    out = ours(*args)
    err = (out.cpu() - ref).abs().max().item()
    assert err <= LOGIT_TOL, err
    assert torch.equal(out, ours(*args))  # fixed summation orders: bit-identical run to run
    ours.fused = False
    batched = ours(*args)
    assert (batched - out).abs().max().item() <= LOGIT_TOL
    # no masks at all (AMO_CLIP.forward(rgb, motion)): every key is valid
    ours.fused = True
    assert (ours(args[0], args[1]).cpu() - oracle(rgb, mot)).abs().max().item() <= LOGIT_TOL


def test_tfam_long_clips_take_the_batched_path(cuda_device):
    """More than 32 frames per clip: outside the fused kernel's shared-memory budget, the batched GEMM path serves it."""
    oracle = otfam.TfamOracle().eval()
    weights.randomise_tfam_(oracle, 6)
    ours = vmc.AMO_CLIP(device=cuda_device)
    ours.load_state_dict(oracle.state_dict(), strict=True)
    ours = ours.to(cuda_device).eval()
    gen = torch.Generator().manual_seed(3)
    rgb, mot = torch.randn(3, 40, 512, generator=gen), torch.randn(3, 39, 512, generator=gen)
    ops.reset_launch_count()
    out = ours(rgb.to(cuda_device), mot.to(cuda_device))
    assert ops.launch_count() > 6
    assert (out.cpu() - oracle(rgb, mot)).abs().max().item() <= LOGIT_TOL


def test_tfam_no_mask_and_larger_batch(cuda_device, golden):
    g = golden("tfam.npz")
    oracle = otfam.TfamOracle().eval()
    weights.randomise_tfam_(oracle, 0)
    ours = vmc.AMO_CLIP(device=cuda_device)
    ours.load_state_dict(oracle.state_dict(), strict=True)
    ours = ours.to(cuda_device).eval()
    out = ours(torch.from_numpy(g["nomask_rgb"]).to(cuda_device), torch.from_numpy(g["nomask_motion"]).to(cuda_device))
    assert (out.cpu() - torch.from_numpy(g["nomask_logits"])).abs().max().item() <= LOGIT_TOL
    # batch 64 x T 16/15 with ragged lengths against the oracle
    gen = torch.Generator().manual_seed(8)
    rgb, mot = torch.randn(64, 16, 512, generator=gen), torch.randn(64, 15, 512, generator=gen)
    lr = torch.randint(1, 17, (64,), generator=gen)
    lm = torch.randint(1, 16, (64,), generator=gen)
    mr = torch.arange(16)[None] < lr[:, None]
    mm = torch.arange(15)[None] < lm[:, None]
    ref = oracle(rgb, mot, mr, mm)
    out = ours(rgb.to(cuda_device), mot.to(cuda_device), mr.to(cuda_device), mm.to(cuda_device))
    assert (out.cpu() - ref).abs().max().item() <= LOGIT_TOL


def test_full_pipeline_small(cuda_device):
    """Config-4 chain in memory on 3 clips: RGB tower + student + TFAM against the oracle chain."""
    torch.manual_seed(0)
    pipe = vmc.ViMoCLIPPipeline("ViT-B/16", "ViT-B/32", device=cuda_device, clips_per_step=2)
    o_rgb = clip_shim.build_visual("ViT-B/16", seed=1)
    o_st = ostudent.StudentOracle("ViT-B/32", seed=2)
    weights.randomise_heads_(o_st, 2)
    o_tf = otfam.TfamOracle().eval()
    weights.randomise_tfam_(o_tf, 3)
    pipe.rgb.visual.load_state_dict(o_rgb.state_dict())
    pipe.student.load_state_dict(o_st.state_dict())
    pipe.tfam.load_state_dict(o_tf.state_dict())
    gen = torch.Generator().manual_seed(21)
    rgb = torch.randint(0, 256, (3, 4, 3, 224, 224), dtype=torch.uint8, generator=gen)
    mot = torch.randint(0, 256, (3, 3, 3, 224, 224), dtype=torch.uint8, generator=gen)
    logits, er, em = pipe(rgb, mot)
    with torch.no_grad():
        er_ref = o_rgb(torch.from_numpy(prologue.normalise_u8(rgb.reshape(12, 3, 224, 224).numpy()))).view(3, 4, -1)
        em_ref, _, _ = o_st(mot)
        lg_ref = o_tf(er_ref, em_ref)  # fp32 reference chain end to end
        lg_stage = o_tf(er.cpu(), em.cpu())  # oracle TFAM on OUR embeddings: stage-3 parity on identical inputs
    assert _cos_min(er, er_ref) >= COS_MIN and _cos_min(em, em_ref) >= COS_MIN
    assert (logits.cpu() - lg_stage).abs().max().item() <= LOGIT_TOL
    # end to end the bf16 embedding error (cos >= 0.9995) feeds the fusion block: looser, stated bound
    assert (logits.cpu() - lg_ref).abs().max().item() <= 5e-2
    assert math.isfinite(float(logits.abs().sum()))


# ---------------------------------------------------------------------------------------------------
# TFAM training step: TFAM/train_and_eval.py:66-101 through our forward + backward kernels
# ---------------------------------------------------------------------------------------------------
def _tfam_pair(cuda_device, dropout=0.0, mlp_dropout=0.0, seed=0, **mode):
    o_tf = otfam.TfamOracle(dropout=dropout, mlp_dropout=mlp_dropout, **mode)
    with torch.no_grad():
        weights.randomise_tfam_(o_tf, seed)
    ours = vmc.AMO_CLIP(dropout=dropout, mlp_dropout=mlp_dropout, device=cuda_device, **mode)
    ours.load_state_dict(o_tf.state_dict(), strict=True)
    return o_tf, ours.to(cuda_device)


def test_backward_building_blocks(cuda_device):
    gen = torch.Generator(device="cuda").manual_seed(3)
    M, K, N = 96, 512, 140
    x = torch.randn(M, K, device=cuda_device, generator=gen, requires_grad=True)
    w = (torch.randn(N, K, device=cuda_device, generator=gen) * K**-0.5).requires_grad_()
    dy = torch.randn(M, N, device=cuda_device, generator=gen)
    from vimoclip_b200.tfam_train import _lin_bwd

    dx, dw, db = _lin_bwd(dy, x.detach(), w.detach())
    (x @ w.t()).backward(dy)
    for got, ref in ((dx, x.grad), (dw, w.grad), (db, dy.sum(0))):
        assert (got - ref).norm().item() <= 2e-4 * ref.norm().item()
    # column sums: single-stage (short) and two-stage (tall) paths, with and without the element-wise product
    for R, Cn in ((96, 140), (5000, 768), (30000, 100)):
        xs = torch.randn(R, Cn, device=cuda_device, generator=gen)
        ys = torch.randn(R, Cn, device=cuda_device, generator=gen)
        assert torch.allclose(ops.colsum(xs), xs.double().sum(0).float(), rtol=1e-4, atol=1e-3)
        assert torch.allclose(ops.colsum(xs, ys), (xs.double() * ys.double()).sum(0).float(), rtol=1e-4, atol=1e-3)
        assert torch.equal(ops.colsum(xs, ys), ops.colsum(xs, ys))  # deterministic
    # LayerNorm backward
    z = (torch.randn(M, K, device=cuda_device, generator=gen) * 2 + 0.3).requires_grad_()
    g_ = (1 + 0.1 * torch.randn(K, device=cuda_device, generator=gen)).requires_grad_()
    b_ = torch.zeros(K, device=cuda_device, requires_grad=True)
    dyl = torch.randn(M, K, device=cuda_device, generator=gen)
    torch.nn.functional.layer_norm(z, (K,), g_, b_, 1e-5).backward(dyl)
    dz, dg, dbt = ops.layernorm_bwd(z.detach(), g_.detach(), 1e-5, dyl)
    for got, ref in ((dz, z.grad), (dg, g_.grad), (dbt, b_.grad)):
        assert (got - ref).norm().item() <= 1e-4 * ref.norm().item()
    # masked attention backward with probability dropout
    B, Tq, Tk, h = 3, 16, 15, 8
    d = h * 64
    q = torch.randn(B * Tq, d, device=cuda_device, generator=gen, requires_grad=True)
    k = torch.randn(B * Tk, d, device=cuda_device, generator=gen, requires_grad=True)
    v = torch.randn(B * Tk, d, device=cuda_device, generator=gen, requires_grad=True)
    valid = torch.ones(B, Tk, dtype=torch.bool, device=cuda_device)
    valid[1, 11:] = False
    pm = ((torch.rand(B, h, Tq, Tk, device=cuda_device, generator=gen) >= 0.2).float() / 0.8).contiguous()
    dO = torch.randn(B * Tq, d, device=cuda_device, generator=gen)
    qh, kh, vh = (t.view(B, -1, h, 64).transpose(1, 2) for t in (q, k, v))
    s = qh @ kh.transpose(-1, -2) / 8.0
    s = s.masked_fill(~valid[:, None, None, :], float("-inf"))
    ref_o = ((torch.softmax(s, -1) * pm) @ vh).transpose(1, 2).reshape(B * Tq, d)
    got_o = ops.attention_masked(q.detach(), k.detach(), v.detach(), valid, B, Tq, Tk, h, out_dtype=torch.float32, prob_mask=pm)
    assert (got_o - ref_o).abs().max().item() < 1e-4
    ref_o.backward(dO)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    ops.attention_masked_bwd(q.detach(), k.detach(), v.detach(), valid, pm, dO, B, Tq, Tk, h, dq, dk, dv)
    for got, ref in ((dq, q.grad), (dk, k.grad), (dv, v.grad)):
        assert (got - ref).norm().item() <= 1e-4 * ref.norm().item()


@pytest.mark.parametrize("tag", list(MODES))
def test_tfam_training_step_gradients_match_reference_autograd(cuda_device, tag):
    """config 1 batch (B=2, T=16/15, ragged masks), dropout off, every fusion mode of the ablation grid: logits and EVERY
    parameter gradient of the BCEWithLogits loss equal fp32 autograd of the reference restatement (oracle pinned to
    TFAM/models/AMO_CLIP.py); parameters the mode does not use get no gradient on either side."""
    o_tf, ours = _tfam_pair(cuda_device, **MODES[tag])
    o_tf.train()
    ours.train()
    gen = torch.Generator().manual_seed(11)
    rgb, mot = torch.randn(2, 16, 512, generator=gen), torch.randn(2, 15, 512, generator=gen)
    m_rgb = torch.arange(16)[None, :] < torch.tensor([16, 12])[:, None]  # collate_fn_pad masks (True = real frame)
    m_mot = torch.arange(15)[None, :] < torch.tensor([15, 11])[:, None]
    labels = (torch.rand(2, 140, generator=gen) < 0.05).float()
    crit = torch.nn.BCEWithLogitsLoss()
    ref_logits = o_tf(rgb.clone(), mot.clone(), m_rgb, m_mot)
    crit(ref_logits, labels).backward()
    logits = ours(rgb.clone().to(cuda_device), mot.clone().to(cuda_device), m_rgb.to(cuda_device), m_mot.to(cuda_device))
    assert logits.requires_grad
    assert (logits.detach().cpu() - ref_logits.detach()).abs().max().item() <= 1e-2
    crit(logits, labels.to(cuda_device)).backward()
    ref_grads = dict(o_tf.named_parameters())
    checked = 0
    for name, p in ours.named_parameters():
        ref = ref_grads[name].grad
        if ref is None:
            assert p.grad is None, name  # unused by this fusion mode, as in the reference
            continue
        err = (p.grad.cpu() - ref).norm().item()
        # 5e-3: pre-activations within ~1e-6 of the ReLU kink flip their gate between the two fp32 summation orders
        assert err <= 5e-3 * ref.norm().item() + 1e-9, (name, err, ref.norm().item())
        checked += 1
    assert checked == {"cross": 78, "cross_pe": 78, "rgb_only": 54, "flow_only": 54, "concat_t": 54, "concat_e": 56}[tag]


def test_tfam_training_loop_with_dropout_reduces_loss(cuda_device):
    """AdamW(lr 1e-4, wd 0.1) as TFAM/train_and_eval.py:53-58 with the reference's dropout rates: the loss goes down."""
    _, ours = _tfam_pair(cuda_device, dropout=0.1, mlp_dropout=0.3, seed=1)
    ours.train()
    gen = torch.Generator().manual_seed(5)
    rgb = torch.randn(32, 16, 512, generator=gen).to(cuda_device)
    mot = torch.randn(32, 15, 512, generator=gen).to(cuda_device)
    labels = (torch.rand(32, 140, generator=gen) < 0.05).float().to(cuda_device)
    opt = torch.optim.AdamW(ours.parameters(), lr=1e-4, weight_decay=0.1)
    crit = torch.nn.BCEWithLogitsLoss()
    losses = []
    for _ in range(12):
        opt.zero_grad()
        loss = crit(ours(rgb, mot), labels)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(l == l for l in losses)
    assert sum(losses[-3:]) < 0.8 * sum(losses[:3]), losses
    ours.eval()
    assert not ours(rgb, mot).requires_grad


# ---------------------------------------------------------------------------------------------------
# Student training step: train.py:86-107 through our forward + backward kernels
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,T", [("ViT-B/32", 3), ("ViT-B/16", 1)])
def test_student_training_step_gradients_match_reference_autograd(cuda_device, name, T):
    """2 clips x T frames through the student in .train() mode (ViT-B/32: the reference's training tower, 50 tokens;
    ViT-B/16: 197 tokens, tiled attention backward): the loss of train.py:98-101 (cosine distillation
    against teacher embeddings + BCE with positive weight) backpropagated through our kernels gives every parameter
    gradient of the tower and the heads to bf16 accuracy against fp32 autograd of the reference restatement."""
    oracle = ostudent.StudentOracle(name, seed=0)
    with torch.no_grad():
        weights.randomise_heads_(oracle, 0)
    ours = vmc.FlowStudentModel(name, device=cuda_device, num_classes=140, alpha=0.1)
    ours.load_state_dict(oracle.state_dict(), strict=True)
    oracle.train()
    ours.train()
    gen = torch.Generator().manual_seed(21)
    frames = torch.randint(0, 256, (2, T, 3, 224, 224), dtype=torch.uint8, generator=gen)
    teacher = torch.randn(2, T, 512, generator=gen)
    labels = (torch.rand(2, 140, generator=gen) < 0.05).float()

    def loss_of(model, dev):
        emb, dis, logits = model(frames)
        l = olosses.distillation_loss(dis, teacher.to(dev), mode="cosine") + olosses.classification_loss(logits, labels.to(dev), 10.0)
        return l + 0.01 * emb.pow(2).mean()  # the raw embeddings get a gradient of their own too

    loss_ref = loss_of(oracle, "cpu")
    loss_ref.backward()
    loss = loss_of(ours, cuda_device)
    assert abs(loss.item() - loss_ref.item()) <= 2e-2 * abs(loss_ref.item())
    loss.backward()
    ref = dict(oracle.named_parameters())
    worst = 0.0
    n = 0
    for name, p in ours.named_parameters():
        r = ref[name].grad
        assert p.grad is not None and r is not None, name
        err = (p.grad.cpu() - r).norm().item() / (r.norm().item() + 1e-12)
        worst = max(worst, err)
        assert err <= 6e-2, (name, err, r.norm().item())
        n += 1
    assert n == 5 + 12 * 12 + 3 + 8
    print(f"student training step: worst relative gradient error {worst:.3e} over {n} tensors")


def test_student_training_loop_reduces_loss(cuda_device):
    """Adam(lr 1e-5) on ALL parameters as train.py:66: the distillation + classification loss goes down."""
    torch.manual_seed(3)
    ours = vmc.FrameDiffStudentModel("ViT-B/32", device=cuda_device, num_classes=140).train()
    gen = torch.Generator().manual_seed(8)
    frames = torch.randint(0, 256, (4, 4, 3, 224, 224), dtype=torch.uint8, generator=gen).to(cuda_device)
    teacher = torch.randn(4, 4, 512, generator=gen).to(cuda_device)
    labels = (torch.rand(4, 140, generator=gen) < 0.05).float().to(cuda_device)
    opt = torch.optim.Adam(ours.parameters(), lr=1e-5)
    losses = []
    for _ in range(8):
        opt.zero_grad()
        emb, dis, logits = ours(frames)
        loss = vmc.losses.distillation_loss(dis, teacher, mode="cosine") + vmc.losses.classification_loss(logits, labels)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(l == l for l in losses) and losses[-1] < losses[0], losses
    ours.eval()
    assert not ours(frames)[0].requires_grad


def test_clip_image_processor_dropin_matches_hf(cuda_device):
    """extract_embeddings.py:89-94: PIL frames -> processor -> .to(device) -> get_image_features(**inputs).  The GPU
    processor (Pillow-exact bicubic resize + centre crop + P1 normalise) against the installed HF CLIPImageProcessor."""
    from PIL import Image
    from transformers import CLIPImageProcessor as HFProcessor

    from torchvision.transforms import CenterCrop, Compose, InterpolationMode, Normalize, Resize, ToTensor

    pil_path = Compose([Resize(224, interpolation=InterpolationMode.BICUBIC), CenterCrop(224), ToTensor(),
                        Normalize(clip_shim.CLIP_MEAN, clip_shim.CLIP_STD)])  # the PIL arithmetic of the pinned 4.53.2 slow path
    gen = np.random.default_rng(12)
    for (H, W) in ((224, 224), (360, 640)):
        imgs = [Image.fromarray(gen.integers(0, 256, (H, W, 3), dtype=np.uint8)) for _ in range(3)]
        inputs = vmc.CLIPImageProcessor()(images=imgs, return_tensors="pt").to(cuda_device)
        got = inputs["pixel_values"].cpu()
        ref_pil = torch.stack([pil_path(im) for im in imgs])
        assert got.shape == ref_pil.shape
        assert torch.equal(got, ref_pil), (H, W)  # bit-exact against the PIL path
        ref = HFProcessor()(images=imgs, return_tensors="pt")["pixel_values"]
        diff = (got - ref).abs()
        if (H, W) == (224, 224):  # no resampling: only the <= 1 ulp rescale/normalise difference
            assert diff.max().item() <= 5e-7
        else:  # the installed transformers 5.5 resizes with torchvision's tensor bicubic: off by one uint8 level in places
            assert diff.max().item() <= 1.05 / 255 / 0.26 and (diff > 1e-6).float().mean().item() < 0.1
    feats = vmc.CLIPVisionFeatures("openai/clip-vit-base-patch32").to(cuda_device)
    out = feats.get_image_features(**inputs)
    assert out.shape == (3, 512) and torch.isfinite(out).all()


@pytest.mark.parametrize("B,Tq,Tk,h,drop", [(2, 150, 197, 4, True), (1, 197, 197, 12, False), (3, 33, 300, 2, True),
                                            (2, 100, 90, 3, True), (2, 64, 64, 2, True), (3, 16, 15, 8, True)])
@pytest.mark.parametrize("path", ["auto", "tiled"])
def test_attention_backward_all_paths(cuda_device, B, Tq, Tk, h, drop, path):
    """Masked attention backward against torch autograd: the register-tiled short-sequence kernel (<= 64 x 64), the
    shared-memory kernel (<= ~128 x 128) and the tiled flash-style kernels (any length; forced with ``tiled``)."""
    gen = torch.Generator(device="cuda").manual_seed(Tq + Tk)
    d = h * 64
    q = torch.randn(B * Tq, d, device=cuda_device, generator=gen, requires_grad=True)
    k = torch.randn(B * Tk, d, device=cuda_device, generator=gen, requires_grad=True)
    v = torch.randn(B * Tk, d, device=cuda_device, generator=gen, requires_grad=True)
    valid = torch.ones(B, Tk, dtype=torch.bool, device=cuda_device)
    valid[0, Tk - Tk // 4:] = False
    pm = ((torch.rand(B, h, Tq, Tk, device=cuda_device, generator=gen) >= 0.1).float() / 0.9).contiguous() if drop else None
    dO = torch.randn(B * Tq, d, device=cuda_device, generator=gen)
    qh, kh, vh = (t.view(B, -1, h, 64).transpose(1, 2) for t in (q, k, v))
    s = (qh @ kh.transpose(-1, -2) / 8.0).masked_fill(~valid[:, None, None, :], float("-inf"))
    p = torch.softmax(s, -1)
    ((p * pm if drop else p) @ vh).transpose(1, 2).reshape(B * Tq, d).backward(dO)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    ops.attention_masked_bwd(q.detach(), k.detach(), v.detach(), valid, pm, dO, B, Tq, Tk, h, dq, dk, dv,
                             tiled=True if path == "tiled" else None)
    for got, ref in ((dq, q.grad), (dk, k.grad), (dv, v.grad)):
        assert (got - ref).norm().item() <= 1e-4 * ref.norm().item()


@pytest.mark.parametrize("impl,tol", [(0, 1e-2), (2, 1e-4)])
@pytest.mark.parametrize("F_,L,h", [(5, 50, 12), (3, 64, 2), (2, 7, 1), (4, 33, 3), (40, 50, 12)])
def test_attention_vit_backward_from_bf16_qkv(cuda_device, F_, L, h, impl, tol):
    """ViT self-attention backward straight from the packed bf16 qkv buffer (ViT-B/32: 50 tokens) against fp32 autograd on the
    cast: impl 0 = warp-level tensor-core kernel (P, dS and dO rounded to bf16 like every GEMM operand of the step),
    impl 2 = register-tiled fp32 kernel."""
    gen = torch.Generator(device="cuda").manual_seed(50 + L)
    d = h * 64
    qkv = torch.randn(F_ * L, 3 * d, device=cuda_device, generator=gen).to(torch.bfloat16)
    dO = torch.randn(F_ * L, d, device=cuda_device, generator=gen)
    ops.set_option(vmc._lib.OPT_ATTN_BWD_IMPL, impl)
    try:
        got = ops.attention_vit_bwd(qkv, dO, F_, L, h)
        again = ops.attention_vit_bwd(qkv, dO, F_, L, h)
    finally:
        ops.set_option(vmc._lib.OPT_ATTN_BWD_IMPL, 0)
    assert torch.equal(got, again)  # deterministic
    q32 = qkv.float().requires_grad_()
    qh, kh, vh = (q32[:, i * d:(i + 1) * d].view(F_, L, h, 64).transpose(1, 2) for i in range(3))
    (torch.softmax(qh @ kh.transpose(-1, -2) / 8.0, -1) @ vh).transpose(1, 2).reshape(F_ * L, d).backward(dO)
    for i, name in enumerate("qkv"):
        a, b = got[:, i * d:(i + 1) * d], q32.grad[:, i * d:(i + 1) * d]
        assert (a - b).norm().item() <= tol * b.norm().item(), name


@pytest.mark.parametrize("M,N,K", [(768, 2304, 6400), (140, 256, 96), (3072, 768, 25600), (500, 333 // 8 * 8 + 8, 1000), (40000, 768, 768)])
@pytest.mark.parametrize("a_t,w_t", [(True, True), (False, True), (True, False)])
def test_gemm_transposed_operands(cuda_device, M, N, K, a_t, w_t):
    """MN-major UMMA operands: A and / or W given transposed in memory ([K, M] / [K, N] row-major), as the backward GEMMs
    dW = dY^T X and dX = dY W read them -- no transposing pass."""
    gen = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = (torch.randn(M, K, device=cuda_device, generator=gen)).to(torch.bfloat16)
    w = (torch.randn(N, K, device=cuda_device, generator=gen) * K**-0.5).to(torch.bfloat16)
    bias = torch.randn(N, device=cuda_device, generator=gen)
    ref = a.float() @ w.float().t() + bias
    def transposed(x):  # [R, C] -> view [C, R] of a buffer whose row stride is a multiple of 8 elements (TMA: 16-byte strides)
        R, Cn = x.shape
        buf = torch.zeros(Cn, (R + 7) // 8 * 8, dtype=x.dtype, device=x.device)
        buf[:, :R] = x.t()
        return buf[:, :R]

    a_arg = transposed(a) if a_t else a
    w_arg = transposed(w) if w_t else w
    got = ops.gemm(a_arg, w_arg, bias=bias, out_dtype=torch.float32, a_t=a_t, w_t=w_t)
    assert got.shape == (M, N)
    assert (got - ref).abs().max().item() < 2e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("ln_mode", [6, 4])
@pytest.mark.parametrize("name,nframes", [("ViT-B/32", 5), ("ViT-B/16", 3), ("ViT-L/14", 2)])
def test_vit_tower_last_block_cls_only(cuda_device, name, nframes, ln_mode):
    """Default (last_block_cls = 1): the last block computes only the CLS row of its attention / MLP output (all the tower
    output reads).  Same embeddings as the full computation (last_block_cls = 2), for the bf16- and fp32-stream towers."""
    torch.manual_seed(1)
    tower = vmc.VisionTower.from_name(name).to(cuda_device)
    tower.ln_mode = ln_mode
    gen = torch.Generator().manual_seed(9)
    u8 = torch.randint(0, 256, (nframes, 3, 224, 224), dtype=torch.uint8, generator=gen).to(cuda_device)
    patches = ops.prologue(u8, wrap=False, dst="patch", patch=tower.patch_size)
    tower.last_block_cls = 2
    full = tower.forward_patches(patches, nframes).clone()
    tower.last_block_cls = 1
    short = tower.forward_patches(patches, nframes).clone()
    tower.last_block_cls = 0
    assert torch.equal(short, tower.forward_patches(patches, nframes))  # 0 = library default = 1
    cos = torch.nn.functional.cosine_similarity(full.double(), short.double(), dim=-1).min().item()
    # the CLS path keeps the probabilities in fp32 (and, with the bf16 stream, the last block's CLS rows) where the full path rounds
    assert cos >= (0.99999 if ln_mode == 4 else 0.9999), cos
    assert (full - short).abs().max().item() <= (5e-3 if ln_mode == 4 else 2e-2) * full.abs().max().item()


# ------------------------------------------------------------------------------------------------
# CUDA-graph replay of the launch-bound regime (inference.py:129: one clip at a time)
# ------------------------------------------------------------------------------------------------
def test_graphed_student_forward_matches_eager(cuda_device):
    torch.manual_seed(3)
    model = vmc.FlowStudentModel("ViT-B/32", device=cuda_device, num_classes=140).eval()
    with torch.no_grad():
        model.residual_mlp.fc2.weight.normal_(0, 0.02)
    gen = torch.Generator().manual_seed(11)
    clips = [torch.randint(0, 256, (1, 15, 3, 224, 224), dtype=torch.uint8, generator=gen).to(cuda_device) for _ in range(3)]
    g = vmc.graphed(model, clips[0])
    for c in clips + clips[:1]:  # replay with new contents, and again with the first clip
        want = [t.clone() for t in model(c)]
        got = g(c)
        for a, b in zip(got, want):
            assert a.shape == b.shape and torch.equal(a, b)
    with pytest.raises(ValueError):
        g(clips[0][:, :7])


def test_graphed_tfam_and_pipeline_match_eager(cuda_device):
    torch.manual_seed(4)
    tfam = vmc.AMO_CLIP(num_classes=140, device=cuda_device).to(cuda_device).eval()
    gen = torch.Generator().manual_seed(12)
    rgb = torch.randn(2, 16, 512, generator=gen).to(cuda_device)
    mot = torch.randn(2, 15, 512, generator=gen).to(cuda_device)
    mask_rgb = (torch.arange(16)[None, :] < torch.tensor([16, 12])[:, None]).to(cuda_device)
    mask_mot = (torch.arange(15)[None, :] < torch.tensor([15, 11])[:, None]).to(cuda_device)
    g = vmc.graphed(tfam, rgb, mot, mask_rgb, mask_mot)
    for scale in (1.0, -0.5):
        want = tfam(rgb * scale, mot * scale, mask_rgb, mask_mot).clone()
        assert torch.equal(g(rgb * scale, mot * scale, mask_rgb, mask_mot), want)
    g2 = vmc.graphed(tfam, rgb, mot, None, None)  # non-tensor arguments are baked in
    assert torch.equal(g2(rgb, mot, None, None), tfam(rgb, mot))
    pipe = vmc.ViMoCLIPPipeline("openai/clip-vit-base-patch32", "ViT-B/32", num_classes=140, device=cuda_device, clips_per_step=1)
    r8 = torch.randint(0, 256, (1, 16, 3, 224, 224), dtype=torch.uint8, generator=gen).to(cuda_device)
    m8 = torch.randint(0, 256, (1, 15, 3, 224, 224), dtype=torch.uint8, generator=gen).to(cuda_device)
    gp = vmc.graphed(pipe, r8, m8)
    want = [t.clone() for t in pipe(r8.flip(1).contiguous(), m8)]
    got = gp(r8.flip(1).contiguous(), m8)
    for a, b in zip(got, want):
        assert torch.equal(a, b)


@pytest.mark.parametrize("R,C,qgelu", [(25600, 768, False), (1000, 3072, True), (37, 96, True), (50, 2304, False), (300, 200, True)])
def test_cast_colsum_fused_pass(cuda_device, R, C, qgelu):
    """vmc_cast_colsum == (QuickGELU backward,) bf16 cast and fp32 column sum done separately."""
    gen = torch.Generator(device="cuda").manual_seed(R + C)
    x = torch.randn(R, C, device=cuda_device, generator=gen)
    aux = torch.randn(R, C, device=cuda_device, generator=gen) * 2 if qgelu else None
    y16, cs = ops.cast_colsum(x, aux)
    v = x.double()
    if qgelu:
        sg = torch.sigmoid(1.702 * aux.double())
        v = v * sg * (1 + 1.702 * aux.double() * (1 - sg))
    assert (y16.double() - v).abs().max().item() <= 2.0 ** -8 * v.abs().max().item() + 1e-6  # one bf16 rounding (+ __expf)
    ref = v.sum(0)
    assert (cs.double() - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item())
    y2, cs2 = ops.cast_colsum(x, aux)
    assert torch.equal(cs, cs2) and torch.equal(y16.view(torch.int16), y2.view(torch.int16))  # deterministic


@pytest.mark.parametrize("rows,d,with_add", [(25600, 768, True), (300, 512, False), (77, 1024, True), (5000, 640, True)])
def test_layernorm_backward_fused_matches_autograd(cuda_device, rows, d, with_add):
    """Fused LayerNorm backward (dz + residual gradient, dgamma / dbeta partials in registers; d = 512 / 768 / 1024) and the generic
    fallback (d = 640) against fp32 autograd."""
    gen = torch.Generator(device="cuda").manual_seed(rows + d)
    z = torch.randn(rows, d, device=cuda_device, generator=gen) * 1.5 + 0.3
    gamma = torch.randn(d, device=cuda_device, generator=gen) * 0.2 + 1.0
    beta = torch.randn(d, device=cuda_device, generator=gen) * 0.1
    dy = torch.randn(rows, d, device=cuda_device, generator=gen)
    add = torch.randn(rows, d, device=cuda_device, generator=gen) if with_add else None
    zz, gg, bb = z.clone().requires_grad_(), gamma.clone().requires_grad_(), beta.clone().requires_grad_()
    torch.nn.functional.layer_norm(zz, (d,), gg, bb, 1e-5).backward(dy)
    want_dz = zz.grad + (add if with_add else 0)
    dz, dg, db = ops.layernorm_bwd(z, gamma, 1e-5, dy, add=add)
    assert (dz - want_dz).abs().max().item() <= 2e-5 * max(1.0, want_dz.abs().max().item())
    assert (dg - gg.grad).abs().max().item() <= 2e-4 * max(1.0, gg.grad.abs().max().item())
    assert (db - bb.grad).abs().max().item() <= 2e-4 * max(1.0, bb.grad.abs().max().item())
    dz2, dg2, db2 = ops.layernorm_bwd(z, gamma, 1e-5, dy, add=add)
    assert torch.equal(dz, dz2) and torch.equal(dg, dg2) and torch.equal(db, db2)  # deterministic


def test_tfam_training_step_as_one_cuda_graph(cuda_device):
    """Forward + BCE loss + backward (our kernels) + AdamW captured in ONE CUDA graph (tools/tfam_train_graph.py): replays must
    produce the same parameters as the same number of eager steps (dropout off: the kernels are deterministic)."""
    gen = torch.Generator().manual_seed(21)
    B = 6
    rgb = torch.randn(B, 16, 512, generator=gen).to(cuda_device)
    mot = torch.randn(B, 15, 512, generator=gen).to(cuda_device)
    m_r = (torch.arange(16)[None, :] < torch.tensor([16, 12, 9, 16, 5, 14])[:, None]).to(cuda_device)
    m_m = (torch.arange(15)[None, :] < torch.tensor([15, 11, 8, 15, 4, 13])[:, None]).to(cuda_device)
    labels = (torch.rand(B, 140, generator=gen) < 0.05).float().to(cuda_device)
    crit = torch.nn.BCEWithLogitsLoss()

    def make():
        torch.manual_seed(5)
        model = vmc.AMO_CLIP(num_classes=140, dropout=0.0, mlp_dropout=0.0, device=cuda_device).to(cuda_device).train()
        return model, torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=0.1, capturable=True)

    def eager(model, opt):
        opt.zero_grad(set_to_none=True)
        loss = crit(model(rgb, mot, m_r, m_m), labels)
        loss.backward()
        opt.step()
        return loss

    ref_model, ref_opt = make()
    for _ in range(5):
        eager(ref_model, ref_opt)
    model, opt = make()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            eager(model, opt)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    opt.zero_grad(set_to_none=True)
    with torch.cuda.graph(g):
        loss = crit(model(rgb, mot, m_r, m_m), labels)
        loss.backward()
        opt.step()
    g.replay()  # capture itself does not execute: 3 eager + 2 replays = 5 steps
    g.replay()
    torch.cuda.synchronize()
    for (n, a), (_, b) in zip(model.named_parameters(), ref_model.named_parameters()):
        assert (a - b).abs().max().item() <= 1e-5 * max(1.0, b.abs().max().item()), n
