"""Freeze golden vectors from THE REFERENCE ITSELF (run in the build container only).

    python oracle/make_golden.py            # writes tests/golden/*.npz

Reads ``/root/reference`` (never copied) plus the reference's third-party numerics that ARE
installed here (HF ``transformers`` CLIP tower, torchvision / PIL transforms, OpenCV).  The GPU
box has no ``/root/reference``; tests there read only the committed ``.npz`` fixtures.

Fixtures (all seeded; weights are regenerated at test time by ``oracle.weights`` and verified with
a checksum stored in the fixture):
  prologue.npz     torchvision ``to_pil_image`` + clip ``_transform`` tables for the 3 input regimes
  framediff.npz    cv2.cvtColor(BGR2GRAY) + cv2.absdiff on random frames (+ exhaustive 2^24 check flag)
  student.npz      reference ``FrameDiffStudentModel`` / ``FlowStudentModel`` files run under the clip shim
  vit_hf.npz       HF ``CLIPVisionModelWithProjection`` with the shim's weights mapped in (B/32, B/16, L/14)
  tfam.npz         reference ``TFAM/models/AMO_CLIP.py`` logits, BASELINE config 1 + every fusion mode
  indexing.npz     reference ``sparse_sampling`` / ``collate_fn_pad`` (TFAM/data/dataset.py)
  losses.npz       reference ``losses.py``
  resize.npz       PIL / torchvision bicubic Resize(224) + CenterCrop(224); reference student on 640x360 frames
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import clip_shim, weights  # noqa: E402


def load_ref_module(name: str, relpath: str):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def config1_tfam_inputs():
    """BASELINE config 1: seeded ragged batch."""
    g = torch.Generator().manual_seed(0)
    rgb = torch.randn(2, 16, 512, generator=g)
    motion = torch.randn(2, 15, 512, generator=g)
    len_rgb = torch.tensor([16, 12])
    len_mot = torch.tensor([15, 11])
    mask_rgb = torch.arange(16)[None, :] < len_rgb[:, None]
    mask_mot = torch.arange(15)[None, :] < len_mot[:, None]
    # collate_fn_pad zero-pads: zero the padded rows like the reference loader does
    rgb = rgb * mask_rgb[..., None]
    motion = motion * mask_mot[..., None]
    return rgb, motion, mask_rgb, mask_mot


def golden_prologue():
    from torchvision.transforms import Compose
    from torchvision.transforms.functional import to_pil_image

    tf = Compose(clip_shim._transform(224).transforms)
    # one 224x224 frame whose pixels run through all 256 values in every channel
    idx = (np.arange(224 * 224).reshape(224, 224)[None] + 37 * np.arange(3)[:, None, None]) % 256
    frame_u8 = torch.from_numpy(idx.astype(np.uint8))
    # regime A: uint8 -> .float() -> to_pil_image (wraps) -> transform
    outA = tf(to_pil_image(frame_u8.float()))
    # regime B: float in [0,1]
    outB = tf(to_pil_image(frame_u8.float() / 255.0))
    # regime C: already-normalised floats (the transform applied twice, inference.py:60-61)
    outC = tf(to_pil_image(outB))
    pilA = np.asarray(to_pil_image(frame_u8.float())).transpose(2, 0, 1)
    pilC = np.asarray(to_pil_image(outB)).transpose(2, 0, 1)
    np.savez_compressed(
        os.path.join(OUT, "prologue.npz"),
        frame_u8=frame_u8.numpy(),
        wrapA_u8=pilA,
        normA=outA.numpy(),
        normB=outB.numpy(),
        wrapC_u8=pilC,
        normC=outC.numpy(),
    )
    print("prologue.npz", outA.shape, float(outA.abs().max()))


def golden_framediff():
    import cv2

    rng = np.random.default_rng(7)
    frames = rng.integers(0, 256, size=(5, 48, 64, 3), dtype=np.uint8)
    gray = np.stack([cv2.cvtColor(f, cv2.COLOR_BGR2GRAY) for f in frames])
    diff = np.stack([cv2.absdiff(gray[t + 1], gray[t]) for t in range(4)])
    # exhaustive check of the closed-form grey formula against OpenCV over all 2^24 colours
    from oracle import prologue

    allc = np.arange(1 << 24, dtype=np.uint32)
    img = np.stack([(allc & 255), (allc >> 8) & 255, (allc >> 16) & 255], -1).astype(np.uint8).reshape(4096, 4096, 3)
    exhaustive_ok = bool(np.array_equal(cv2.cvtColor(img, cv2.COLOR_BGR2GRAY), prologue.bgr2gray(img)))
    np.savez_compressed(os.path.join(OUT, "framediff.npz"), frames=frames, gray=gray, diff=diff, exhaustive_ok=exhaustive_ok)
    print("framediff.npz exhaustive 2^24 colours match:", exhaustive_ok)


def golden_student():
    clip_shim.install()
    sys.path.insert(0, REF)
    out = {}
    for tag, relpath, cls, name, shape in [
        ("fd_b32", "models/student_model_frame_diff.py", "FrameDiffStudentModel", "ViT-B/32", (2, 16)),
        ("flow_b32", "models/student_model.py", "FlowStudentModel", "ViT-B/32", (1, 3)),
    ]:
        mod = load_ref_module("ref_" + tag, relpath)
        clip_shim.set_seed(0)
        model = getattr(mod, cls)(name, device="cpu", num_classes=140, alpha=0.1)
        weights.randomise_heads_(model, 0)
        model.eval()
        g = torch.Generator().manual_seed(1234)
        frames = torch.randint(0, 256, (*shape, 3, 224, 224), dtype=torch.uint8, generator=g)
        with torch.no_grad():
            emb, dis, logits = model(frames)
        out[tag + "_emb"] = emb.numpy()
        out[tag + "_distill"] = dis.numpy()
        out[tag + "_logits"] = logits.numpy()
        out[tag + "_checksum"] = np.float64(weights.state_checksum(model.state_dict()))
        print(tag, emb.shape, float(emb.abs().mean()), float(logits.abs().max()))
    # float-input regimes through the reference forward (regime B: [0,1]; regime C: normalised)
    mod = load_ref_module("ref_fd_small", "models/student_model_frame_diff.py")
    clip_shim.set_seed(0)
    model = mod.FrameDiffStudentModel("ViT-B/32", device="cpu", num_classes=140)
    weights.randomise_heads_(model, 0)
    model.eval()
    g = torch.Generator().manual_seed(99)
    u8 = torch.randint(0, 256, (1, 2, 3, 224, 224), dtype=torch.uint8, generator=g)
    with torch.no_grad():
        embB, _, logB = model(u8.float() / 255.0)
        normed = (u8.float() / 255.0 - torch.tensor(clip_shim.CLIP_MEAN).view(1, 1, 3, 1, 1)) / torch.tensor(clip_shim.CLIP_STD).view(1, 1, 3, 1, 1)
        embC, _, logC = model(normed)
    out.update(regB_emb=embB.numpy(), regB_logits=logB.numpy(), regC_emb=embC.numpy(), regC_logits=logC.numpy())
    np.savez_compressed(os.path.join(OUT, "student.npz"), **out)


def hf_from_openai(vit, name):
    """Map OpenAI-layout weights into the installed HF tower (SURVEY.md Appendix A mapping)."""
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection

    patch, width, layers, heads, out_dim = clip_shim.VIT_CONFIGS[name]
    cfg = CLIPVisionConfig(
        hidden_size=width, intermediate_size=4 * width, num_hidden_layers=layers, num_attention_heads=heads,
        patch_size=patch, image_size=224, projection_dim=out_dim, hidden_act="quick_gelu", layer_norm_eps=1e-5,
    )
    hf = CLIPVisionModelWithProjection(cfg).eval()
    sd = vit.state_dict()
    new = {}
    p = "vision_model."
    new[p + "embeddings.patch_embedding.weight"] = sd["conv1.weight"]
    new[p + "embeddings.class_embedding"] = sd["class_embedding"]
    new[p + "embeddings.position_embedding.weight"] = sd["positional_embedding"]
    new[p + "pre_layrnorm.weight"] = sd["ln_pre.weight"]
    new[p + "pre_layrnorm.bias"] = sd["ln_pre.bias"]
    new[p + "post_layernorm.weight"] = sd["ln_post.weight"]
    new[p + "post_layernorm.bias"] = sd["ln_post.bias"]
    new["visual_projection.weight"] = sd["proj"].t().contiguous()
    for i in range(layers):
        s = f"transformer.resblocks.{i}."
        t = p + f"encoder.layers.{i}."
        w, b = sd[s + "attn.in_proj_weight"], sd[s + "attn.in_proj_bias"]
        for j, nm in enumerate(["q_proj", "k_proj", "v_proj"]):
            new[t + f"self_attn.{nm}.weight"] = w[j * width : (j + 1) * width]
            new[t + f"self_attn.{nm}.bias"] = b[j * width : (j + 1) * width]
        new[t + "self_attn.out_proj.weight"] = sd[s + "attn.out_proj.weight"]
        new[t + "self_attn.out_proj.bias"] = sd[s + "attn.out_proj.bias"]
        for a, b_ in [("ln_1", "layer_norm1"), ("ln_2", "layer_norm2"), ("mlp.c_fc", "mlp.fc1"), ("mlp.c_proj", "mlp.fc2")]:
            new[t + b_ + ".weight"] = sd[s + a + ".weight"]
            new[t + b_ + ".bias"] = sd[s + a + ".bias"]
    missing, unexpected = hf.load_state_dict(new, strict=False)
    missing = [m for m in missing if "position_ids" not in m]
    assert not missing and not unexpected, (missing, unexpected)
    return hf


def golden_vit_hf():
    out = {}
    for name, nframes in [("ViT-B/32", 3), ("ViT-B/16", 2), ("ViT-L/14", 1)]:
        tag = name.replace("/", "").replace("-", "").lower()
        vit = clip_shim.build_visual(name, seed=0)
        hf = hf_from_openai(vit, name)
        g = torch.Generator().manual_seed(4321)
        u8 = torch.randint(0, 256, (nframes, 3, 224, 224), dtype=torch.uint8, generator=g)
        from oracle import prologue

        x = torch.from_numpy(prologue.normalise_u8(u8.numpy()))
        with torch.no_grad():
            y_hf = hf(pixel_values=x).image_embeds
            y_shim = vit(x)
        out[tag + "_hf"] = y_hf.numpy()
        out[tag + "_checksum"] = np.float64(weights.state_checksum(vit.state_dict()))
        print(name, "HF vs shim max-abs", float((y_hf - y_shim).abs().max()), "cos", float(torch.cosine_similarity(y_hf, y_shim).min()))
        del vit, hf
    np.savez_compressed(os.path.join(OUT, "vit_hf.npz"), **out)


def golden_tfam():
    ref = load_ref_module("ref_amo", "TFAM/models/AMO_CLIP.py")
    out = {}
    rgb, motion, mask_rgb, mask_mot = config1_tfam_inputs()
    out.update(rgb=rgb.numpy(), motion=motion.numpy(), mask_rgb=mask_rgb.numpy(), mask_mot=mask_mot.numpy())
    modes = {
        "cross": dict(),
        "cross_pe": dict(use_pe=True),
        "rgb_only": dict(use_only_rgb=True),
        "flow_only": dict(use_only_flow=True),
        "concat_t": dict(use_cross_attention=False, concat_dim=1),
        "concat_e": dict(use_cross_attention=False, concat_dim=-1),
    }
    for tag, kw in modes.items():
        torch.manual_seed(0)
        model = ref.AMO_CLIP(device="cpu", **kw).eval()
        weights.randomise_tfam_(model, 0)
        with torch.no_grad():
            logits = model(rgb.clone(), motion.clone(), mask_rgb, mask_mot)
        out[tag + "_logits"] = logits.numpy()
        out[tag + "_checksum"] = np.float64(weights.state_checksum(model.state_dict()))
        print("tfam", tag, logits.shape, float(logits.abs().max()))
    # no-mask call (allowed in the cross-attention branch) with T_m = 16
    torch.manual_seed(0)
    model = ref.AMO_CLIP(device="cpu").eval()
    weights.randomise_tfam_(model, 0)
    g = torch.Generator().manual_seed(5)
    rgb2, mot2 = torch.randn(3, 16, 512, generator=g), torch.randn(3, 16, 512, generator=g)
    with torch.no_grad():
        out["nomask_logits"] = model(rgb2.clone(), mot2.clone()).numpy()
    out["nomask_rgb"], out["nomask_motion"] = rgb2.numpy(), mot2.numpy()
    np.savez_compressed(os.path.join(OUT, "tfam.npz"), **out)


def golden_indexing():
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))  # dataset.py imports h5py at module level only
    ds = load_ref_module("ref_tfam_ds", "TFAM/data/dataset.py")
    out = {}
    cases = [(450, 16), (17, 16), (16, 16), (10, 16), (1000, 32), (33, 32), (97, 8)]
    for T, n in cases:
        emb = torch.arange(T, dtype=torch.float32)[:, None].repeat(1, 2)
        out[f"sparse_{T}_{n}"] = ds.sparse_sampling(emb, n)[:, 0].long().numpy()
    g = torch.Generator().manual_seed(3)
    lens_r, lens_f = [16, 12, 7], [15, 11, 6]
    batch = [
        {"video_id": str(i), "embeddings": torch.randn(a, 8, generator=g), "flow_embeddings": torch.randn(b, 8, generator=g), "labels": torch.zeros(4)}
        for i, (a, b) in enumerate(zip(lens_r, lens_f))
    ]
    col = ds.collate_fn_pad(batch)
    out["collate_rgb"] = col["embeddings"].numpy()
    out["collate_flow"] = col["flow_embeddings"].numpy()
    out["collate_mask_rgb"] = col["mask_rgb"].numpy()
    out["collate_mask_flow"] = col["mask_flow"].numpy()
    for i, b in enumerate(batch):
        out[f"collate_in_rgb{i}"] = b["embeddings"].numpy()
        out[f"collate_in_flow{i}"] = b["flow_embeddings"].numpy()
    np.savez_compressed(os.path.join(OUT, "indexing.npz"), **out)
    print("indexing.npz", len(out), "arrays")


def golden_resize():
    """PIL / torchvision Resize(224, BICUBIC) + CenterCrop(224) on seeded frames, and the reference student file on
    640x360 frames (wrap -> resize -> crop -> normalise -> ViT)."""
    from PIL import Image
    from torchvision.transforms import CenterCrop, Compose, InterpolationMode, Resize

    tf = Compose([Resize(224, interpolation=InterpolationMode.BICUBIC), CenterCrop(224)])
    out = {}
    # the last four: frames SMALLER than 224 (Resize scales the short side UP to 224; CenterCrop then never pads)
    sizes = {"360x640": (360, 640), "240x320": (240, 320), "500x375": (500, 375), "120x160": (120, 160), "100x300": (100, 300),
             "223x225": (223, 225), "64x64": (64, 64)}
    for tag, (H, W) in sizes.items():
        img = np.random.default_rng(len(tag) * 7 + H).integers(0, 256, size=(H, W, 3), dtype=np.uint8)
        out["pil_" + tag] = np.asarray(tf(Image.fromarray(img))).transpose(2, 0, 1)
        out["seed_" + tag] = np.int64(len(tag) * 7 + H)
    # HF CLIPImageProcessor (extract_embeddings.py:18,91): same resampler, crop offset (dim - 224) // 2 -- differs from
    # torchvision's round-half-to-even when the margin is 3 mod 4 (360x648 -> 224x403: margin 179 -> 89 vs 90; 227x224 ->
    # margin 3 -> 1 vs 2).  Frozen from the installed processor as the uint8 crop (pixel_values de-normalised exactly).
    # The reference pins transformers 4.53.2, whose CLIPImageProcessor is the PIL ("slow") processor; the installed 5.5 keeps
    # that implementation as CLIPImageProcessorPil (its default class resizes with torchvision on tensors instead).
    from transformers.models.clip import CLIPImageProcessorPil

    proc = CLIPImageProcessorPil()
    mean = np.array(clip_shim.CLIP_MEAN, dtype=np.float64)[:, None, None]
    std = np.array(clip_shim.CLIP_STD, dtype=np.float64)[:, None, None]
    for tag, (H, W) in {"360x648": (360, 648), "227x224": (227, 224)}.items():
        img = np.random.default_rng(len(tag) * 11 + W).integers(0, 256, size=(H, W, 3), dtype=np.uint8)
        pv = proc(images=[Image.fromarray(img)], return_tensors="np")["pixel_values"][0].astype(np.float64)
        u8 = np.rint((pv * std + mean) * 255.0)
        assert np.abs((pv * std + mean) * 255.0 - u8).max() < 1e-3
        u8 = u8.astype(np.uint8)
        tv = np.asarray(tf(Image.fromarray(img))).transpose(2, 0, 1)
        assert not np.array_equal(u8, tv), "these geometries are chosen so that HF's and torchvision's crops differ"
        out["hf_" + tag] = u8
        out["hfseed_" + tag] = np.int64(len(tag) * 11 + W)
    clip_shim.install()
    sys.path.insert(0, REF)
    mod = load_ref_module("ref_fd_resize", "models/student_model_frame_diff.py")
    clip_shim.set_seed(0)
    model = mod.FrameDiffStudentModel("ViT-B/32", device="cpu", num_classes=140)
    weights.randomise_heads_(model, 0)
    model.eval()
    g = torch.Generator().manual_seed(31)
    frames = torch.randint(0, 256, (1, 2, 3, 360, 640), dtype=torch.uint8, generator=g)
    with torch.no_grad():
        emb, dis, logits = model(frames)
    out.update(student_emb=emb.numpy(), student_logits=logits.numpy())
    np.savez_compressed(os.path.join(OUT, "resize.npz"), **out)
    print("resize.npz", {k: getattr(v, "shape", v) for k, v in out.items()})


def golden_losses():
    ref = load_ref_module("ref_losses", "losses.py")
    g = torch.Generator().manual_seed(11)
    s, t = torch.randn(8, 10, 512, generator=g), torch.randn(8, 10, 512, generator=g)
    t2 = s + 0.1 * torch.randn(8, 10, 512, generator=g)
    logits = torch.randn(8, 140, generator=g)
    targets = (torch.rand(8, 140, generator=g) > 0.9).float()
    np.savez_compressed(
        os.path.join(OUT, "losses.npz"),
        s=s.numpy(), t=t.numpy(), t2=t2.numpy(), logits=logits.numpy(), targets=targets.numpy(),
        cos=ref.distillation_loss(s, t, "cosine").numpy(), cos2=ref.distillation_loss(s, t2, "cosine").numpy(),
        mse=ref.distillation_loss(s, t, "mse").numpy(),
        bce=ref.classification_loss(logits, targets).numpy(), bce_pw=ref.classification_loss(logits, targets, positive_weight=3).numpy(),
    )
    print("losses.npz")


def golden_sampling():
    """extract_embeddings.py:77-81 (frame sampling) -- the module cannot be imported (module-level ``from_pretrained`` needs the
    network; decord / h5py are absent), so the `if (max_frames is None) or ...: indices = ... else: ...` statement is cut
    out of the reference SOURCE with ``ast`` and executed on its own."""
    import ast

    src = open(os.path.join(REF, "extract_embeddings.py")).read()
    tree = ast.parse(src)
    stmt = None
    for node in ast.walk(tree):
        if isinstance(node, ast.If) and "max_frames" in ast.unparse(node.test) and "total_frames" in ast.unparse(node.test):
            targets = {t.id for b in (node.body, node.orelse) for st in b if isinstance(st, ast.Assign) for t in st.targets if isinstance(t, ast.Name)}
            if "indices" in targets:
                stmt = node
                break
    assert stmt is not None, "frame-sampling statement not found in extract_embeddings.py"
    assert 77 <= stmt.lineno <= 78, stmt.lineno  # the lines SURVEY.md section 8 a4 cites
    code = compile(ast.Module(body=[stmt], type_ignores=[]), "extract_embeddings.py:77-81", "exec")
    cases = [(t, m) for t in (1, 5, 15, 16, 17, 31, 32, 33, 47, 64, 100, 250, 1000) for m in (None, 1, 8, 16, 32)]
    tot, mx, rows = [], [], []
    width = max(t for t, _ in cases)
    for t, m in cases:
        ns = {"np": np, "total_frames": t, "max_frames": m}
        exec(code, ns)
        idx = np.asarray(ns["indices"], dtype=np.int64)
        row = np.full(width, -1, dtype=np.int64)
        row[: len(idx)] = idx
        tot.append(t)
        mx.append(-1 if m is None else m)
        rows.append(row)
    np.savez_compressed(os.path.join(OUT, "sampling.npz"), total_frames=np.array(tot), max_frames=np.array(mx),
                        indices_flat_split=np.stack(rows), source_lines=np.array([stmt.lineno, stmt.end_lineno]))
    print("sampling.npz", len(cases), "cases from lines", stmt.lineno, stmt.end_lineno)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_grad_enabled(False)
    which = sys.argv[1:] or ["prologue", "framediff", "student", "vit_hf", "tfam", "indexing", "losses", "resize", "sampling"]
    for w in which:
        globals()["golden_" + w]()
