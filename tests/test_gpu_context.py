"""Context numbers: the reference's own modules in stock PyTorch eager on the SAME B200 -- fp32 as the
reference runs them (`model.float()`, models/student_model.py:45) and under bf16 autocast -- next to this repository's
path on identical inputs.  The oracle restatement stands in for the reference modules (pinned to them by the golden
fixtures); its CPU preprocessing loop is left out, so the eager numbers are an UPPER bound for the reference."""
import time

import pytest
import torch

import vimoclip_b200 as vmc
from oracle import clip_shim, prologue, student as ostudent, tfam as otfam, weights


@pytest.mark.gpu
def test_report_torch_eager_context(cuda_device):
    clips, T, Tm = 128, 16, 15  # one chunk of the bench step (2048 RGB frames in flight)
    gen = torch.Generator().manual_seed(1234)
    rgb = torch.randint(0, 256, (clips, T, 3, 224, 224), dtype=torch.uint8, generator=gen)
    mot = torch.randint(0, 256, (clips, Tm, 3, 224, 224), dtype=torch.uint8, generator=gen)
    rgb_tower = clip_shim.build_visual("ViT-B/16", seed=0).to(cuda_device).eval()
    student = ostudent.StudentOracle("ViT-B/32", seed=0)
    weights.randomise_heads_(student, 0)
    student = student.to(cuda_device).eval()
    tfam = otfam.TfamOracle().eval()
    weights.randomise_tfam_(tfam, 0)
    tfam = tfam.to(cuda_device)
    x_rgb = torch.from_numpy(prologue.normalise_u8(rgb.reshape(-1, 3, 224, 224).numpy())).to(cuda_device)
    x_mot = torch.from_numpy(prologue.preprocess_frames(mot.reshape(-1, 3, 224, 224).numpy())).to(cuda_device)

    def eager(dtype):
        with torch.no_grad(), torch.autocast("cuda", dtype=dtype, enabled=dtype != torch.float32):
            er = rgb_tower(x_rgb).float().view(clips, T, -1)
            em = student.visual_encoder(x_mot).float().view(clips, Tm, -1)
            return tfam(er, em)

    pipe = vmc.ViMoCLIPPipeline("openai/clip-vit-base-patch16", "ViT-B/32", num_classes=140, device=cuda_device)
    pipe.rgb.visual.load_state_dict(rgb_tower.state_dict(), strict=True)
    pipe.student.load_state_dict(student.state_dict(), strict=True)
    pipe.tfam.load_state_dict(tfam.state_dict(), strict=True)
    rgb_d, mot_d = rgb.to(cuda_device), mot.to(cuda_device)

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            out = fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps, out

    t32, ref = timed(lambda: eager(torch.float32))
    t16, out16 = timed(lambda: eager(torch.bfloat16))
    tours, (logits, _, _) = timed(lambda: pipe(rgb_d, mot_d))
    fps = lambda t: clips * T / t  # noqa: E731
    err_ours = (logits - ref).abs().max().item()
    err_16 = (out16.float() - ref).abs().max().item()
    print(f"\ncontext ({clips} clips, preprocessing excluded for eager): torch eager fp32 {fps(t32):.0f} frames/s, torch eager bf16 autocast "
          f"{fps(t16):.0f} frames/s, this repository (uint8 frames in, prologue included) {fps(tours):.0f} frames/s; logit max-abs vs eager "
          f"fp32: ours {err_ours:.2e}, eager bf16 autocast {err_16:.2e}")
    assert err_ours <= 1e-2
    assert tours < t16
    # latency regime: one clip, as inference.py:129 (batch size 1) feeds the student
    x1 = x_mot[:Tm]
    m1 = mot_d[:1]

    def eager_student(dtype):
        with torch.no_grad(), torch.autocast("cuda", dtype=dtype, enabled=dtype != torch.float32):
            return student.visual_encoder(x1)

    l32, _ = timed(lambda: eager_student(torch.float32), reps=10)
    l16, _ = timed(lambda: eager_student(torch.bfloat16), reps=10)
    lours, _ = timed(lambda: pipe.student(m1), reps=10)
    print(f"latency, ONE clip of {Tm} frames through the ViT-B/32 student: torch eager fp32 {l32 * 1e3:.2f} ms, eager bf16 {l16 * 1e3:.2f} ms, "
          f"this repository {lours * 1e3:.2f} ms")


@pytest.mark.gpu
def test_report_training_step_context(cuda_device):
    """Student training step (train.py:86-107) on 32 clips x 16 frames (the reference trains on sequences of hundreds of
    frames, models/student_model.py:104): stock PyTorch eager autograd of the reference modules
    (fp32 and bf16 autocast) next to this repository's forward + backward kernels, same weights and inputs."""
    from oracle import losses as olosses

    clips, T = 32, 16
    gen = torch.Generator().manual_seed(7)
    frames = torch.randint(0, 256, (clips, T, 3, 224, 224), dtype=torch.uint8, generator=gen)
    teacher = torch.randn(clips, T, 512, generator=gen).to(cuda_device)
    labels = (torch.rand(clips, 140, generator=gen) < 0.05).float().to(cuda_device)
    oracle = ostudent.StudentOracle("ViT-B/32", seed=0)
    weights.randomise_heads_(oracle, 0)
    oracle = oracle.to(cuda_device).train()
    ours = vmc.FlowStudentModel("ViT-B/32", device=cuda_device, num_classes=140)
    ours.load_state_dict(oracle.state_dict(), strict=True)
    ours.train()
    x = torch.from_numpy(prologue.preprocess_frames(frames.reshape(-1, 3, 224, 224).numpy())).to(cuda_device)
    frames_d = frames.to(cuda_device)

    def eager_step(dtype):
        for p in oracle.parameters():
            p.grad = None
        with torch.autocast("cuda", dtype=dtype, enabled=dtype != torch.float32):
            emb = oracle.visual_encoder(x).float().view(clips, T, -1)
            dis = oracle.residual_mlp(emb)
            logits = oracle.classification_head(emb.mean(dim=1))
        loss = olosses.distillation_loss(dis, teacher, mode="cosine") + olosses.classification_loss(logits, labels)
        loss.backward()
        return loss

    def our_step():
        for p in ours.parameters():
            p.grad = None
        emb, dis, logits = ours(frames_d)
        loss = vmc.losses.distillation_loss(dis, teacher, mode="cosine") + vmc.losses.classification_loss(logits, labels)
        loss.backward()
        return loss

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            out = fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps, out

    t32, l32 = timed(lambda: eager_step(torch.float32))
    t16, _ = timed(lambda: eager_step(torch.bfloat16))
    tours, lours = timed(our_step)
    print(f"\nstudent training step, {clips} clips x {T} frames (ViT-B/32, forward + backward, no optimiser): torch eager fp32 {t32 * 1e3:.1f} ms, "
          f"eager bf16 autocast {t16 * 1e3:.1f} ms, this repository {tours * 1e3:.1f} ms; loss {l32.item():.5f} vs {lours.item():.5f}")
    assert abs(l32.item() - lours.item()) <= 2e-2 * abs(l32.item())
