"""CPU: the oracle restatement against the golden vectors frozen from the reference itself
(oracle/make_golden.py ran the reference's own files / HF / torchvision / OpenCV in the build container)."""
import numpy as np
import torch

from oracle import clip_shim, indexing as oidx, losses as olosses, prologue, student as ostudent, tfam as otfam, weights


def test_prologue_regimes_bit_exact(golden):
    g = golden("prologue.npz")
    u8 = g["frame_u8"][None]  # [1,3,224,224]
    # regime A: uint8 0..255 -> (-x) mod 256
    assert np.array_equal(prologue.to_pil_u8(u8)[0], g["wrapA_u8"])
    assert np.array_equal(prologue.to_pil_u8(u8)[0], (-u8[0].astype(np.int32) % 256).astype(np.uint8))
    assert np.array_equal(prologue.preprocess_frames(u8)[0].view(np.uint32), g["normA"].view(np.uint32))
    # regime B: float [0,1] round-trips exactly
    fB = u8.astype(np.float32) / np.float32(255)
    assert np.array_equal(prologue.to_pil_u8(fB), u8)
    assert np.array_equal(prologue.preprocess_frames(fB)[0].view(np.uint32), g["normB"].view(np.uint32))
    # regime C: already-normalised floats (transform applied twice)
    fC = g["normB"][None]
    assert np.array_equal(prologue.to_pil_u8(fC)[0], g["wrapC_u8"])
    assert np.array_equal(prologue.preprocess_frames(fC)[0].view(np.uint32), g["normC"].view(np.uint32))


def test_known_wrap_values():
    # SURVEY.md Appendix B.1 probes
    x = np.array([-1.7923, -0.5, 0.3, 1.0, 2.1459], dtype=np.float32)
    assert prologue.to_pil_u8(x).tolist() == [55, 129, 76, 255, 35]
    assert prologue.to_pil_u8(np.array([0, 1, 2, 255], dtype=np.uint8)).tolist() == [0, 255, 254, 1]


def test_frame_difference_matches_opencv(golden):
    g = golden("framediff.npz")
    assert bool(g["exhaustive_ok"])  # all 2^24 colours checked against cv2 when the fixture was made
    assert np.array_equal(prologue.bgr2gray(g["frames"]), g["gray"])
    assert np.array_equal(prologue.frame_difference(g["frames"]), g["diff"])


def test_patchify_matches_conv_weight_order():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 3, 64, 64)).astype(np.float32)
    w = rng.standard_normal((5, 3, 16, 16)).astype(np.float32)
    ref = torch.nn.functional.conv2d(torch.from_numpy(x), torch.from_numpy(w), stride=16)  # [2,5,4,4]
    ref = ref.reshape(2, 5, 16).permute(0, 2, 1).reshape(32, 5).numpy()
    got = prologue.patchify(x, 16) @ w.reshape(5, -1).T
    assert np.allclose(got, ref, atol=1e-4)


def _student(name="ViT-B/32"):
    m = ostudent.StudentOracle(name, seed=0)
    weights.randomise_heads_(m, 0)
    return m.eval()


def test_student_oracle_matches_reference_files(golden):
    g = golden("student.npz")
    m = _student()
    assert abs(weights.state_checksum(m.state_dict()) - float(g["fd_b32_checksum"])) < 1e-6 * abs(float(g["fd_b32_checksum"]))
    gen = torch.Generator().manual_seed(1234)
    frames = torch.randint(0, 256, (2, 16, 3, 224, 224), dtype=torch.uint8, generator=gen)
    emb, dis, logits = m(frames)
    assert np.allclose(emb.numpy(), g["fd_b32_emb"], atol=2e-5)
    assert np.allclose(dis.numpy(), g["fd_b32_distill"], atol=2e-5)
    assert np.allclose(logits.numpy(), g["fd_b32_logits"], atol=2e-5)
    gen = torch.Generator().manual_seed(1234)
    frames = torch.randint(0, 256, (1, 3, 3, 224, 224), dtype=torch.uint8, generator=gen)
    emb, dis, logits = m(frames)
    assert np.allclose(emb.numpy(), g["flow_b32_emb"], atol=2e-5)
    assert np.allclose(logits.numpy(), g["flow_b32_logits"], atol=2e-5)


def test_student_float_regimes(golden):
    g = golden("student.npz")
    m = _student()
    gen = torch.Generator().manual_seed(99)
    u8 = torch.randint(0, 256, (1, 2, 3, 224, 224), dtype=torch.uint8, generator=gen)
    embB, _, logB = m(u8.float() / 255.0)
    assert np.allclose(embB.numpy(), g["regB_emb"], atol=2e-5) and np.allclose(logB.numpy(), g["regB_logits"], atol=2e-5)
    mean = torch.tensor(clip_shim.CLIP_MEAN).view(1, 1, 3, 1, 1)
    std = torch.tensor(clip_shim.CLIP_STD).view(1, 1, 3, 1, 1)
    embC, _, logC = m((u8.float() / 255.0 - mean) / std)
    assert np.allclose(embC.numpy(), g["regC_emb"], atol=2e-5) and np.allclose(logC.numpy(), g["regC_logits"], atol=2e-5)


def test_vit_restatement_matches_hf_tower(golden):
    g = golden("vit_hf.npz")
    for name, nframes in [("ViT-B/32", 3), ("ViT-B/16", 2)]:
        tag = name.replace("/", "").replace("-", "").lower()
        vit = clip_shim.build_visual(name, seed=0)
        ck = float(g[tag + "_checksum"])
        assert abs(weights.state_checksum(vit.state_dict()) - ck) < 1e-6 * abs(ck)
        gen = torch.Generator().manual_seed(4321)
        u8 = torch.randint(0, 256, (nframes, 3, 224, 224), dtype=torch.uint8, generator=gen)
        with torch.no_grad():
            y = vit(torch.from_numpy(prologue.normalise_u8(u8.numpy())))
        assert np.allclose(y.numpy(), g[tag + "_hf"], atol=2e-5), name


def _tfam_inputs(g):
    return (torch.from_numpy(g["rgb"]), torch.from_numpy(g["motion"]), torch.from_numpy(g["mask_rgb"]), torch.from_numpy(g["mask_mot"]))


MODES = {
    "cross": dict(),
    "cross_pe": dict(use_pe=True),
    "rgb_only": dict(use_only_rgb=True),
    "flow_only": dict(use_only_flow=True),
    "concat_t": dict(use_cross_attention=False, concat_dim=1),
    "concat_e": dict(use_cross_attention=False, concat_dim=-1),
}


def test_tfam_oracle_matches_reference_file(golden):
    g = golden("tfam.npz")
    rgb, mot, mr, mm = _tfam_inputs(g)
    for tag, kw in MODES.items():
        m = otfam.TfamOracle(**kw).eval()
        weights.randomise_tfam_(m, 0)
        ck = float(g[tag + "_checksum"])
        assert abs(weights.state_checksum(m.state_dict()) - ck) < 1e-6 * abs(ck), tag
        out = m(rgb.clone(), mot.clone(), mr, mm)
        assert np.allclose(out.numpy(), g[tag + "_logits"], atol=3e-5), tag
    m = otfam.TfamOracle().eval()
    weights.randomise_tfam_(m, 0)
    out = m(torch.from_numpy(g["nomask_rgb"]), torch.from_numpy(g["nomask_motion"]))
    assert np.allclose(out.numpy(), g["nomask_logits"], atol=3e-5)


def test_indexing_matches_reference(golden):
    g = golden("indexing.npz")
    for key in g.files:
        if key.startswith("sparse_"):
            _, T, n = key.split("_")
            assert np.array_equal(oidx.sparse_sampling_indices(int(T), int(n)).numpy(), g[key]), key
    seqs_r = [torch.from_numpy(g[f"collate_in_rgb{i}"]) for i in range(3)]
    seqs_f = [torch.from_numpy(g[f"collate_in_flow{i}"]) for i in range(3)]
    pr, mr = oidx.pad_and_mask(seqs_r)
    pf, mf = oidx.pad_and_mask(seqs_f)
    assert np.array_equal(pr.numpy(), g["collate_rgb"]) and np.array_equal(mr.numpy(), g["collate_mask_rgb"])
    assert np.array_equal(pf.numpy(), g["collate_flow"]) and np.array_equal(mf.numpy(), g["collate_mask_flow"])
    # extract_embeddings.py:77-81 frame sampling
    assert oidx.sample_frame_indices(10, None).tolist() == list(range(10))
    assert oidx.sample_frame_indices(10, 16).tolist() == list(range(10))
    assert oidx.sample_frame_indices(100, 16).tolist() == list(range(0, 100, 6))[:16]
    assert oidx.segment_indices(5, 3) == [[0, 1, 2], [3, 4, 4]]


def test_losses_match_reference(golden):
    g = golden("losses.npz")
    s, t, t2 = (torch.from_numpy(g[k]) for k in ("s", "t", "t2"))
    assert np.isclose(float(olosses.distillation_loss(s, t, "cosine")), float(g["cos"]), atol=1e-6)
    assert np.isclose(float(olosses.distillation_loss(s, t2, "cosine")), float(g["cos2"]), atol=1e-6)
    assert np.isclose(float(olosses.distillation_loss(s, t, "mse")), float(g["mse"]), atol=1e-6)
    lg, tg = torch.from_numpy(g["logits"]), torch.from_numpy(g["targets"])
    assert np.isclose(float(olosses.classification_loss(lg, tg)), float(g["bce"]), atol=1e-6)
    assert np.isclose(float(olosses.classification_loss(lg, tg, positive_weight=3)), float(g["bce_pw"]), atol=1e-6)


def test_resize_oracle_matches_pil(golden):
    from oracle import resize

    g = golden("resize.npz")
    # the last four are SMALLER than 224: Resize scales the short side up, CenterCrop never pads
    for tag, (H, W) in {"360x640": (360, 640), "240x320": (240, 320), "500x375": (500, 375), "120x160": (120, 160),
                        "100x300": (100, 300), "223x225": (223, 225), "64x64": (64, 64)}.items():
        img = np.random.default_rng(int(g["seed_" + tag])).integers(0, 256, size=(H, W, 3), dtype=np.uint8)
        got = resize.resize_center_crop_u8(np.ascontiguousarray(img.transpose(2, 0, 1)))
        assert np.array_equal(got, g["pil_" + tag]), tag
    # HF CLIPImageProcessor (extract_embeddings.py:18,91): same resampler, floor crop offset
    for tag, (H, W) in {"360x648": (360, 648), "227x224": (227, 224)}.items():
        img = np.random.default_rng(int(g["hfseed_" + tag])).integers(0, 256, size=(H, W, 3), dtype=np.uint8)
        chw = np.ascontiguousarray(img.transpose(2, 0, 1))
        assert np.array_equal(resize.resize_center_crop_u8(chw, hf_crop=True), g["hf_" + tag]), tag
        assert not np.array_equal(resize.resize_center_crop_u8(chw), g["hf_" + tag]), tag
    assert resize.resized_size(360, 640) == (224, 398) and resize.resized_size(640, 360) == (398, 224)
    # same-size frames are untouched (PIL's identity shortcut), and bicubic at scale 1 is an exact identity
    x = np.random.default_rng(0).integers(0, 256, size=(3, 224, 224), dtype=np.uint8)
    assert np.array_equal(resize.resize_center_crop_u8(x), x)
    b, k = resize.precompute_coeffs(50, 50)
    assert all(k[i, i - b[i, 0]] == 1 << 22 for i in range(50))


def test_student_oracle_on_640x360_frames(golden):
    g = golden("resize.npz")
    m = _student()
    gen = torch.Generator().manual_seed(31)
    frames = torch.randint(0, 256, (1, 2, 3, 360, 640), dtype=torch.uint8, generator=gen)
    emb, _, logits = m(frames)
    assert np.allclose(emb.numpy(), g["student_emb"], atol=2e-5) and np.allclose(logits.numpy(), g["student_logits"], atol=2e-5)


def test_frame_sampling_indices_match_reference_source(golden):
    """extract_embeddings.py:77-81 cannot be imported (module-level from_pretrained), so oracle/make_golden.py cut the
    statement out of the reference SOURCE with `ast`, executed it and froze the results: the oracle and the product's host
    function reproduce them bit for bit."""
    import numpy as np

    from oracle import indexing as oidx
    from vimoclip_b200 import indexing

    g = golden("sampling.npz")
    assert list(g["source_lines"]) == [77, 81]
    for total, mx, want in zip(g["total_frames"], g["max_frames"], g["indices_flat_split"]):
        want = want[want >= 0]
        m = None if mx < 0 else int(mx)
        assert np.array_equal(np.asarray(oidx.sample_frame_indices(int(total), m)), want), (total, mx)
        assert np.array_equal(indexing.sample_frame_indices(int(total), m), want), (total, mx)
