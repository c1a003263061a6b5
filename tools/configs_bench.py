"""Throughput of the BASELINE.json configurations that are NOT the bench.py headline (config 4): one JSON line per config.

    python tools/configs_bench.py [--configs 2,3,3t,5] [--steps K] [--warmup W]
    python -m torch.distributed.run --nproc-per-node 8 ... tools/configs_bench.py --configs 5      # clip-sharded + NCCL gather

  2   CLIP ViT-B/16 per-frame embedding extraction (extract_embeddings.py:89-94), 16-frame 224x224 uint8 clips, frames in flight swept
  3   MoCLIP frame-difference student: BGR clips [N,17,224,224,3] -> fused frame-diff prologue -> ViT-B/32 student + heads -> cosine
      distillation loss against CLIP ViT-B/16 teacher embeddings [:, :-1] (train_frame_diff.py / train.py:98), inference kernels
  3t  the same as a TRAINING step (student .train(): forward + backward through the whole ViT-B/32 + Adam over all parameters)
  5   CLIP ViT-L/14 at 32 frames/clip (extract_embeddings_mammalNet.py:47-55), clip-sharded, all-gather of [clips*32, 768] fp32

Same timing rules as bench.py (W >= 3 warm-up steps, CUDA events on the launching stream, barrier + synchronize on both sides, max over
ranks, inputs resident in HBM and larger than L2).  The per-class times come from the library's event profiler (one extra step).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vimoclip_b200 as vmc  # noqa: E402
from vimoclip_b200 import _lib, ops  # noqa: E402
from bench import load_peaks  # noqa: E402

FLOPS = {"ViT-B/32": 8.818e9, "ViT-B/16": 35.127e9, "ViT-L/14": 162.03e9}  # SURVEY.md 8d, per frame


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="2,3,3t,5")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--clips", type=int, default=0, help="clips per GPU per step (0 = per-config default)")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    L = _lib.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / args.steps

    def classes(fn):
        L.vmc_profile_begin()
        fn()
        n = 6
        ms_c, fl_c, by_c, la_c = (C.c_double * n)(), (C.c_double * n)(), (C.c_double * n)(), (C.c_longlong * n)()
        _lib.check(L.vmc_profile_end(ms_c, fl_c, by_c, la_c, n), "vmc_profile_end")
        names = ["prologue", "gemm_tcgen05", "attention_vit", "layernorm", "attention_tfam", "other"]
        return {names[i]: {"ms": round(ms_c[i], 3), "launches": la_c[i], "tflops": round(fl_c[i] / (ms_c[i] * 1e9), 1) if ms_c[i] > 0 and fl_c[i] > 0 else None,
                           "gbs": round(by_c[i] / (ms_c[i] * 1e6), 1) if ms_c[i] > 0 and by_c[i] > 0 else None} for i in range(n) if la_c[i]}

    def emit(config, ms, frames, flops, extra):
        if rank == 0:
            tf = flops / (ms * 1e9)
            print(json.dumps({"config": config, "n_gpus": world, "frames_per_s": world * frames / (ms / 1e3), "ms_per_step": ms, "steps": args.steps,
                              "warmup": args.warmup, "tflops_per_gpu": tf, "frac_of_bf16_sustained": tf / peaks["bf16_sustained"],
                              "frac_of_bf16_burst": tf / peaks["bf16_burst"], "dtype": "bf16", "data": "synthetic", **extra}), flush=True)

    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    for cfg in args.configs.split(","):
        torch.manual_seed(0)
        if cfg == "2":
            clip = vmc.CLIPVisionFeatures("openai/clip-vit-base-patch16").to(dev)
            for clips in ([args.clips] if args.clips else [16, 64, 256]):
                frames = torch.randint(0, 256, (clips * 16, 3, 224, 224), dtype=torch.uint8, device=dev, generator=gen)
                clip.visual.frames_in_flight = min(2048, clips * 16)
                fn = lambda: clip.get_image_features_u8(frames)  # noqa: E731
                ms = timed(fn)
                emit("config2: CLIP ViT-B/16 per-frame embedding extraction, 16-frame 224x224 uint8 clips", ms, clips * 16, clips * 16 * FLOPS["ViT-B/16"],
                     {"clips_per_gpu": clips, "frames_in_flight": clip.visual.frames_in_flight, "kernel_classes": classes(fn)})
            del clip, frames
        elif cfg in ("3", "3t"):
            clips = args.clips or (128 if cfg == "3" else 32)
            teacher = vmc.CLIPVisionFeatures("openai/clip-vit-base-patch16").to(dev)
            student = vmc.FrameDiffStudentModel("ViT-B/32", device=dev, num_classes=140)
            with torch.no_grad():
                student.residual_mlp.fc2.weight.normal_(0, 0.02)
            bgr = torch.randint(0, 256, (clips, 17, 224, 224, 3), dtype=torch.uint8, device=dev, generator=gen)
            rgb = bgr.flip(-1).permute(0, 1, 4, 2, 3).contiguous()
            labels = (torch.rand(clips, 140, device=dev, generator=gen) < 0.02).float()
            with torch.no_grad():  # teacher embeddings are precomputed by extract_embeddings.py in the reference (HDF5); sliced as train.py:98
                emb_gt = teacher.get_image_features_u8(rgb.view(-1, 3, 224, 224)).view(clips, 17, -1)[:, :-1, :].contiguous()
            if cfg == "3":
                student.eval()

                def fn():
                    with torch.no_grad():
                        _, e_distill, logits = student.forward_bgr(bgr)
                        return vmc.distillation_loss(e_distill, emb_gt, "cosine")
                ms = timed(fn)
                emit("config3: MoCLIP frame-difference student (uint8 BGR frame-diff prologue + ViT-B/32 + heads) + cosine distillation loss vs CLIP teacher, inference kernels",
                     ms, clips * 16, clips * 16 * FLOPS["ViT-B/32"], {"clips_per_gpu": clips, "kernel_classes": classes(fn)})
            else:
                student.train()
                opt = torch.optim.Adam(student.parameters(), lr=1e-5)
                # the reference trains on decoded frame-difference videos (uint8 [B,T,3,H,W], dataset.py:98-99): materialise them once
                with torch.no_grad():
                    d = ops.frame_diff(bgr, dst=None, want_diff=True)[0]
                    videos = d.view(clips, 16, 1, 224, 224).expand(-1, -1, 3, -1, -1).contiguous()

                def fn():
                    opt.zero_grad(set_to_none=True)
                    _, e_distill, logits = student(videos)
                    loss = vmc.distillation_loss(e_distill, emb_gt, "cosine") + vmc.classification_loss(logits, labels)
                    loss.backward()
                    opt.step()
                    return loss
                ms = timed(fn)
                emit("config3 (training step): frame-difference student forward + backward through ViT-B/32 + heads, cosine distillation + BCE, Adam on all parameters",
                     ms, clips * 16, 3 * clips * 16 * FLOPS["ViT-B/32"], {"clips_per_gpu": clips, "flops_note": "3x forward FLOPs (forward + dX + dW)", "kernel_classes": classes(fn)})
            del teacher, student
        elif cfg == "5":
            clips = args.clips or 64
            clip = vmc.CLIPVisionFeatures("openai/clip-vit-large-patch14").to(dev)
            clip.visual.frames_in_flight = 1024
            frames = torch.randint(0, 256, (clips * 32, 3, 224, 224), dtype=torch.uint8, device=dev, generator=gen)
            total = clips * world

            def fn():
                e = clip.get_image_features_u8(frames).view(clips, 32, -1)
                return vmc.sharding.gather_clips(e, total) if world > 1 else e
            ms = timed(fn)
            emit("config5: CLIP ViT-L/14 frame encoder, 32 frames/clip, clip-sharded, NCCL all-gather of the [clips*32, 768] fp32 embeddings in the step",
                 ms, clips * 32, clips * 32 * FLOPS["ViT-L/14"],
                 {"clips_per_gpu": clips, "sweep_scale": f"{total} of MammalNet's 20033 clips per step", "gather_bytes": total * 32 * 768 * 4, "kernel_classes": classes(fn)})
            del clip, frames
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
