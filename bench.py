"""Benchmark of the ViMoCLIP per-frame encoding hot path (BASELINE.json metric: frames/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--clips C]

Workload (BASELINE config 4, "Full ViMoCLIP inference"): per GPU and per step, C clips (default 256) of
16 RGB frames + 15 motion frames, 224x224 uint8, through CLIP ViT-B/16 (RGB) + MoCLIP ViT-B/32 student
(motion) + TFAM cross-attention fusion -> logits [C,140]; synthetic data, seeded random weights.
One JSON line is printed by rank 0.  `value` = RGB frames/s with inputs resident in HBM; `e2e` = the same
metric through the public `ViMoCLIPPipeline.forward` with pinned HOST inputs (H2D of every frame and D2H of
logits + embeddings inside the timed region).  For N > 1 (torchrun, one rank per GPU) clips are sharded,
every rank runs the same per-GPU batch (weak scaling), and the timed step ends with the NCCL all-gather
of logits and embeddings.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_RGB, T_MOT, RES, NUM_CLASSES = 16, 15, 224, 140
# SURVEY.md section 8d / BASELINE.md section 3: algorithmic FLOPs (2*MAC)
FLOPS_B16, FLOPS_B32 = 35.127e9, 8.818e9
FLOPS_TFAM_CLIP = 0.5371e9  # T_rgb 16, T_mot 15
FLOPS_HEADS_CLIP = 0.0171e9


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_burst": p["bf16_tflops"], "bf16_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "hbm": p["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        # median over the samples taken under load (upper half)
        sm_sorted = sorted(sm)
        under_load = sm_sorted[len(sm_sorted) // 2:]
        return {"sm_mhz": statistics.median(under_load), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own CPU implementation of the path on the box's host cores
# ----------------------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def _load_reference_modules():
    """The reference's own files (staged unmodified by __graft_entry__.build() into baseline/_ref):
    TFAM/models/AMO_CLIP.py as is, models/student_model.py under the `clip` shim (OpenAI clip is not installed anywhere: the
    shim restates clip.load / VisionTransformer, oracle/clip_shim.py).  Returns (AMO_CLIP, FlowStudentModel) or None."""
    import importlib.util

    paths = [os.path.join(REF_DIR, "TFAM", "models", "AMO_CLIP.py"), os.path.join(REF_DIR, "models", "student_model.py")]
    if not all(os.path.exists(p) for p in paths):
        return None
    from oracle import clip_shim

    clip_shim.install()
    mods = []
    for i, p in enumerate(paths):
        spec = importlib.util.spec_from_file_location(f"_vimoclip_reference_{i}", p)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mods.append(mod)
    return mods[0].AMO_CLIP, mods[1].FlowStudentModel


def cpu_reference_run(steps: int, warmup: int, clips: int, kind: str = "auto"):
    """One step = `clips` clips of the bench workload through the reference's CPU path, all host threads.

    kind "reference": stage 1 as extract_embeddings.py:87-94 does it (per-frame PIL images -> HF CLIPImageProcessor -> HF
    CLIPModel.get_image_features; the installed transformers, seeded random ViT-B/16 weights), stage 2 = the reference's
    FlowStudentModel file (its per-frame to_pil_image + preprocess loop included, models/student_model.py:74-81), stage 3 =
    the reference's AMO_CLIP file.  kind "port": the oracle restatement (vectorised numpy preprocessing; faster than the
    reference).  "auto": reference when baseline/_ref is staged, else port."""
    import numpy as np
    import torch

    from oracle import clip_shim, prologue, student as ostudent, tfam as otfam, weights

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ref = _load_reference_modules() if kind in ("auto", "reference") else None
    kind = "reference" if ref is not None else "port"
    gen = torch.Generator().manual_seed(1234)
    rgb = torch.randint(0, 256, (clips, T_RGB, 3, RES, RES), dtype=torch.uint8, generator=gen)
    mot = torch.randint(0, 256, (clips, T_MOT, 3, RES, RES), dtype=torch.uint8, generator=gen)
    if kind == "reference":
        from PIL import Image
        from transformers import CLIPConfig, CLIPModel, CLIPVisionConfig

        try:  # the pinned transformers 4.53.2 processor is the PIL one; v5 keeps it under this name
            from transformers.models.clip import CLIPImageProcessorPil as CLIPImageProcessor
        except ImportError:
            from transformers import CLIPImageProcessor

        AMO_CLIP, FlowStudentModel = ref
        torch.manual_seed(0)
        vcfg = CLIPVisionConfig(hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12, patch_size=16,
                                image_size=224, projection_dim=512)
        clip_model = CLIPModel(CLIPConfig(vision_config=vcfg.to_dict(), projection_dim=512)).eval()
        clip_processor = CLIPImageProcessor()
        student = FlowStudentModel("ViT-B/32", device="cpu", num_classes=NUM_CLASSES).eval()
        tfam = AMO_CLIP(num_classes=NUM_CLASSES, device="cpu").eval()

        def step():
            with torch.no_grad():
                embs = []
                for c in range(clips):  # one video at a time, as the reference's loop does (extract_embeddings.py:61-94)
                    pil_images = [Image.fromarray(np.transpose(frame.numpy(), (1, 2, 0))) for frame in rgb[c]]
                    inputs = clip_processor(images=pil_images, return_tensors="pt")
                    feats = clip_model.get_image_features(inputs["pixel_values"])
                    embs.append(feats if torch.is_tensor(feats) else feats.pooler_output)  # v4.53 returns a tensor, v5 an output object
                er = torch.stack(embs)
                em, _, _ = student(mot)
                return tfam(er, em)
    else:
        rgb_tower = clip_shim.build_visual("ViT-B/16", seed=0)
        student = ostudent.StudentOracle("ViT-B/32", seed=0)
        weights.randomise_heads_(student, 0)
        tfam = otfam.TfamOracle().eval()
        weights.randomise_tfam_(tfam, 0)

        def step():
            with torch.no_grad():
                x = torch.from_numpy(prologue.normalise_u8(rgb.reshape(-1, 3, RES, RES).numpy()))
                er = rgb_tower(x).view(clips, T_RGB, -1)
                em, _, _ = student(mot)
                return tfam(er, em)

    for _ in range(warmup):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    fps = clips * T_RGB / (ms / 1e3)
    return fps, ms, cores, torch.get_num_threads(), kind


def _cpu_sample_text(kind, clips, threads, cores):
    what = ("the reference's own files (baseline/_ref: TFAM/models/AMO_CLIP.py as is, models/student_model.py under the clip shim with its "
            "per-frame PIL loop, HF CLIPImageProcessor + CLIPModel.get_image_features as extract_embeddings.py:87-94)" if kind == "reference"
            else "oracle restatement of the reference (vectorised numpy preprocessing instead of the reference's per-frame PIL loop)")
    return (f"{clips} clip(s) x ({T_RGB} RGB + {T_MOT} motion) frames per step of the same workload; {what}; fp32 PyTorch CPU, "
            f"{threads} threads of {cores} cores")


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # bounded sample: ~0.6-1 s of CPU work per clip on 16 threads; keep the whole --steps/--warmup run within ~3 minutes
    clips = max(1, min(args.ref_clips, int(180 / max(1, args.steps + max(args.warmup, 1)))))
    fps, ms, cores, threads, kind = cpu_reference_run(args.steps, max(args.warmup, 1), clips)
    line = {
        "impl": "reference", "metric": "frames/sec (CLIP ViT+MoCLIP+TFAM)", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.clips),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind, "sample": _cpu_sample_text(kind, clips, threads, cores)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def workload_config(clips):
    return {"workload": f"config4: full ViMoCLIP inference (CLIP ViT-B/16 on RGB + MoCLIP ViT-B/32 student on motion + TFAM), {clips} clips x "
                        f"({T_RGB} RGB + {T_MOT} motion) frames x {RES}x{RES} uint8 per GPU per step, {NUM_CLASSES} classes",
            "clips_per_gpu": clips, "frames_per_clip": T_RGB, "motion_frames_per_clip": T_MOT, "parallelism": "clip-sharded, weights replicated",
            "l2": "inputs (1.19 GB/step) and activations (>8 GB) are larger than the 126 MB L2"}


def other_configs(vmc, ops, dev, rank, world, timed, peaks):
    """Short measurements of BASELINE configs 2, 3 and 5 (config 1 is the CPU parity case, config 4 the headline): one dict
    each, per-GPU batch fixed (weak scaling), `frames_per_s` aggregated over the N GPUs.  tools/configs_bench.py is the
    long form (sweeps, training step, per-class times)."""
    import torch

    FL = {"ViT-B/32": 8.818e9, "ViT-B/16": 35.127e9, "ViT-L/14": 162.03e9}
    gen = torch.Generator(device=dev).manual_seed(4321 + rank)
    out = []

    def rec(name, ms, frames, flops, extra):
        tf = flops / (ms * 1e9)
        out.append({"config": name, "frames_per_s": world * frames / (ms / 1e3), "ms_per_step": ms, "steps": 3, "warmup": 3, "tflops_per_gpu": tf,
                    "frac_of_bf16_sustained": tf / peaks["bf16_sustained"], **extra})

    torch.manual_seed(0)
    # config 2: CLIP ViT-B/16 per-frame embedding extraction (extract_embeddings.py:89-94), 256 clips x 16 frames of uint8
    clip = vmc.CLIPVisionFeatures("openai/clip-vit-base-patch16").to(dev)
    frames = torch.randint(0, 256, (256 * 16, 3, RES, RES), dtype=torch.uint8, device=dev, generator=gen)
    ms = timed(lambda: clip.get_image_features_u8(frames), 3, 3)
    rec("config2: CLIP ViT-B/16 per-frame embedding extraction, 256 clips x 16 frames x 224x224 uint8 per GPU", ms, 256 * 16, 256 * 16 * FL["ViT-B/16"],
        {"clips_per_gpu": 256})
    # config 3: BGR clips -> fused frame-difference prologue -> ViT-B/32 student + heads -> cosine distillation loss vs the teacher [:, :-1]
    student = vmc.FrameDiffStudentModel("ViT-B/32", device=dev, num_classes=NUM_CLASSES).eval()
    bgr = torch.randint(0, 256, (128, 17, RES, RES, 3), dtype=torch.uint8, device=dev, generator=gen)
    with torch.no_grad():  # teacher embeddings are precomputed (HDF5) in the reference; sliced as train.py:98
        emb_gt = clip.get_image_features_u8(frames[: 128 * 17]).view(128, 17, -1)[:, :-1, :].contiguous()

    def fn3():
        with torch.no_grad():
            _, e_distill, _ = student.forward_bgr(bgr)
            return vmc.distillation_loss(e_distill, emb_gt, "cosine")

    ms = timed(fn3, 3, 3)
    rec("config3: MoCLIP frame-difference student (uint8 BGR frame-diff prologue + ViT-B/32 + heads) + cosine distillation loss vs CLIP teacher, 128 clips x 16 "
        "difference frames per GPU", ms, 128 * 16, 128 * 16 * FL["ViT-B/32"], {"clips_per_gpu": 128})
    del clip, student, frames, bgr, emb_gt
    torch.cuda.empty_cache()
    # config 5: CLIP ViT-L/14 at 32 frames per clip, clip-sharded, NCCL all-gather of the [clips * 32, 768] fp32 embeddings in the step
    big = vmc.CLIPVisionFeatures("openai/clip-vit-large-patch14").to(dev)
    big.visual.frames_in_flight = 1024
    clips5 = 64
    frames = torch.randint(0, 256, (clips5 * 32, 3, RES, RES), dtype=torch.uint8, device=dev, generator=gen)

    def fn5():
        e = big.get_image_features_u8(frames).view(clips5, 32, -1)
        return vmc.sharding.gather_clips(e, clips5 * world) if world > 1 else e

    ms = timed(fn5, 3, 3)
    rec("config5: CLIP ViT-L/14 frame encoder, 32 frames per clip, 64 clips per GPU, NCCL all-gather of the embeddings in the step", ms, clips5 * 32,
        clips5 * 32 * FL["ViT-L/14"], {"clips_per_gpu": clips5, "sweep_scale": f"{clips5 * world} of MammalNet's 20033 clips per step",
                                       "gather_bytes": clips5 * world * 32 * 768 * 4})
    del big, frames
    torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------
def ours_main(args):
    import torch
    import torch.distributed as dist

    import vimoclip_b200 as vmc
    from vimoclip_b200 import _lib, ops

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clips = args.clips
    ops.set_option(_lib.OPT_ATTN_IMPL, args.attn_impl)
    ops.set_option(_lib.OPT_PROLOGUE_IMPL, args.prologue_impl)
    torch.manual_seed(0)
    pipe = vmc.ViMoCLIPPipeline("openai/clip-vit-base-patch16", "ViT-B/32", num_classes=NUM_CLASSES, device=dev, clips_per_step=args.chunk)
    for tower in (pipe.rgb.visual, pipe.student.visual_encoder):
        tower.frames_in_flight = args.frames_in_flight
        tower.ln_mode, tower.last_block_cls = args.ln_mode, args.last_block_cls  # per-model selectors (0 = library default)
    pipe.tfam.fused = not args.tfam_batched
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    rgb_dev = torch.randint(0, 256, (clips, T_RGB, 3, RES, RES), dtype=torch.uint8, device=dev, generator=gen)
    mot_dev = torch.randint(0, 256, (clips, T_MOT, 3, RES, RES), dtype=torch.uint8, device=dev, generator=gen)
    total_clips = clips * world

    def step_resident():
        lg, er, em = pipe(rgb_dev, mot_dev)
        if world > 1:
            lg = vmc.sharding.gather_clips(lg, total_clips)
            er = vmc.sharding.gather_clips(er, total_clips)
            em = vmc.sharding.gather_clips(em, total_clips)
        return lg, er, em

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps

    # ---- device-resident throughput (`value`) ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ops.reset_launch_count()
    ms_step = timed(step_resident, args.steps, args.warmup)
    launches = ops.launch_count() // (args.steps + args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    fps = total_clips * T_RGB / (ms_step / 1e3)

    # ---- end to end through the public API with pinned host inputs (`e2e`) ----
    rgb_host = torch.empty(rgb_dev.shape, dtype=torch.uint8).pin_memory()
    mot_host = torch.empty(mot_dev.shape, dtype=torch.uint8).pin_memory()
    rgb_host.copy_(rgb_dev)
    mot_host.copy_(mot_dev)
    out_host = None

    def step_e2e():
        nonlocal out_host
        lg, er, em = pipe(rgb_host, mot_host)  # H2D of every frame inside the call
        if world > 1:
            lg = vmc.sharding.gather_clips(lg, total_clips)
            er = vmc.sharding.gather_clips(er, total_clips)
            em = vmc.sharding.gather_clips(em, total_clips)
        out_host = (lg.cpu(), er[:clips].cpu() if world > 1 else er.cpu(), em[:clips].cpu() if world > 1 else em.cpu())

    e2e_steps = max(1, min(args.steps, 5))
    ms_e2e = timed(step_e2e, e2e_steps, 1)
    fps_e2e = total_clips * T_RGB / (ms_e2e / 1e3)
    h2d = rgb_host.numel() + mot_host.numel()
    d2h = sum(t.numel() * t.element_size() for t in out_host)

    # ---- strong scaling of BASELINE config 4 as it is written: ONE batch of 256 clips, clip-sharded over the N GPUs ----
    strong = None
    if not args.no_strong:
        sc = args.strong_clips
        n_local = len(vmc.indexing.shard_ids(sc, rank, world))
        if n_local > clips:
            raise SystemExit("--strong-clips needs at most --clips clips per GPU")
        rgb_s, mot_s = rgb_dev[:n_local], mot_dev[:n_local]

        def step_strong():
            return pipe.forward_sharded(rgb_s, mot_s, sc)  # towers + TFAM on the local clips, three NCCL all-gathers

        ops.reset_launch_count()
        ms_strong = timed(step_strong, args.steps, 3)
        strong = {"clips_total": sc, "clips_per_gpu": n_local, "value": sc * T_RGB / (ms_strong / 1e3), "unit": "frames/s", "ms_per_step": ms_strong,
                  "scaling": "strong", "gpu_launches_per_step": ops.launch_count() // (args.steps + 3),
                  "note": "ViMoCLIPPipeline.forward_sharded: rank r owns clips r::N; all_gather_into_tensor of logits and both embedding sets inside the step"}

    # ---- the same step with the round-1 defaults, in the same process (N = 1): what the new defaults buy ----
    ab = None
    if world == 1 and not args.no_ab:
        def with_variant(ln_mode, cls, fused):
            for tower in (pipe.rgb.visual, pipe.student.visual_encoder):
                tower.ln_mode, tower.last_block_cls = ln_mode, cls
            pipe.tfam.fused = fused
            return timed(step_resident, 3, 3)

        ab = {"full_last_block_ms": with_variant(args.ln_mode, 2, not args.tfam_batched),
              "round1_defaults_ms": with_variant(4, 2, False),
              "note": "full_last_block: last_block_cls = 2 (every token of the last block computed, identical embeddings); round1_defaults: fp32 residual "
                      "stream + separate LayerNorm kernels + full last block + batched TFAM; 3 timed steps each after 3 warm-ups"}
        with_variant(args.ln_mode, args.last_block_cls, not args.tfam_batched)

    # ---- per-kernel-class profile of one extra step (roofline of the dominant kernel) ----
    L = _lib.lib()
    L.vmc_profile_begin()
    step_resident()
    n = 6
    ms_c, fl_c, by_c, la_c = (C.c_double * n)(), (C.c_double * n)(), (C.c_double * n)(), (C.c_longlong * n)()
    _lib.check(L.vmc_profile_end(ms_c, fl_c, by_c, la_c, n), "vmc_profile_end")
    names = ["prologue", "gemm_tcgen05", "attention_vit", "layernorm", "attention_tfam", "other"]
    classes = {names[i]: {"ms": ms_c[i], "launches": la_c[i], "tflops": (fl_c[i] / (ms_c[i] * 1e9) if ms_c[i] > 0 else None),
                          "gbs": (by_c[i] / (ms_c[i] * 1e6) if ms_c[i] > 0 else None)} for i in range(n)}
    peaks = load_peaks()
    gemm = classes["gemm_tcgen05"]
    # DRAM traffic per launch of the GEMM class: dram__bytes_read + dram__bytes_write summed over the launches of one step
    # in the committed ncu launch list (same command line), divided by the launch count -- ncu cannot run inside the bench
    traffic, traffic_src = None, None
    default_variant = args.ln_mode in (0, 6) and args.last_block_cls in (0, 1) and not args.tfam_batched
    tpath = os.path.join(ROOT, "profiles", "r02_step_launches_dram.json" if default_variant else "r01_step_launches_dram.json")
    round1_variant = args.ln_mode == 4 and args.last_block_cls == 2
    if os.path.exists(tpath) and clips == 256 and args.frames_in_flight == 2048 and (default_variant or round1_variant):
        with open(tpath) as f:
            tj = json.load(f)
        traffic = tj["gemm_class"]["dram_bytes_per_launch"]
        traffic_src = ("STATIC: %s -- ncu dram__bytes_read.sum + dram__bytes_write.sum over the %d GEMM launches of one step of this command line, "
                       "captured once and committed (ncu cannot run inside the timed bench)" % (os.path.relpath(tpath, ROOT), tj["gemm_class"]["launches"]))
    # HBM-bound kernels as the un-bracketed ncu launch list of the same step saw them (static: committed capture).  The event
    # bracketing of the live profile costs a 150 us kernel ~30 % (two event records + the launch gap), ncu times the kernel alone.
    hbm_static = None
    if traffic is not None and default_variant:
        hbm_static = {}
        for kname, label in (("prologue_gather_kernel<16", "prologue_vit_b16"), ("prologue_gather_kernel<32", "prologue_vit_b32"),
                             ("layernorm_kernel", "layernorm")):
            ks = [k for k in tj["kernels"] if k["kernel"].startswith(kname)]
            if ks:
                by = sum(k["dram_read_MB"] + k["dram_write_MB"] for k in ks) * 1e6
                us = sum(k["total_us"] for k in ks)
                hbm_static[label] = {"gbs": by / us / 1e3, "frac_of_hbm_peak": by / us / 1e3 / peaks["hbm"], "launches": sum(k["launches"] for k in ks),
                                     "basis": "DRAM bytes (ncu dram__bytes_read + write) / gpu__time_duration"}
    roofline = {"bound": "tensor", "kernel": "gemm2_bf16_tcgen05_kernel", "achieved": gemm["tflops"], "peak": peaks["bf16_sustained"],
                "unit": "TFLOP/s", "frac": (gemm["tflops"] / peaks["bf16_sustained"]) if gemm["tflops"] else None, "traffic": traffic,
                "traffic_unit": "bytes per launch (DRAM, ncu)", "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": (gemm["gbs"] * gemm["ms"] * 1e6 / max(1, gemm["launches"])) if gemm["gbs"] else None,
                "peak_source": peaks["source"] + ", sustained bf16 (kernel timed inside a long step)",
                "launches_per_step": gemm["launches"], "avg_launch_ms": gemm["ms"] / max(1, gemm["launches"]),
                "share_of_step": gemm["ms"] / sum(c["ms"] for c in classes.values()),
                "hbm_kernels": {k: {"gbs": classes[k]["gbs"], "frac_of_hbm_peak": (classes[k]["gbs"] / peaks["hbm"]) if classes[k]["gbs"] else None}
                                for k in ("prologue", "layernorm")},
                "hbm_kernels_ncu_static": hbm_static}
    # algorithmic FLOPs of the REFERENCE's work per step (every token of every block); the default path skips the dead part of
    # the last block (only its CLS row is read), so the executed FLOPs -- the sum over the launches -- are lower
    flops_step = clips * (T_RGB * FLOPS_B16 + T_MOT * FLOPS_B32 + FLOPS_TFAM_CLIP + FLOPS_HEADS_CLIP)
    step_tflops = flops_step / (ms_step * 1e9)
    executed_flops = sum(fl_c[i] for i in range(n))

    # ---- the other BASELINE configurations, short (configs 2, 3, 5; same timing rules) ----
    other = None
    if not args.no_configs:
        del rgb_host, mot_host
        other = other_configs(vmc, ops, dev, rank, world, timed, peaks)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only) ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cfps, cms, cores, threads, kind = cpu_reference_run(1, 1, args.ref_clips)
        cpu = {"value": cfps, "unit": "frames/s", "cores": threads, "kind": kind,
               "sample": _cpu_sample_text(kind, args.ref_clips, threads, cores) + f"; 1 warm-up + 1 timed pass ({cms / 1e3:.1f} s)"}
        if kind == "reference":  # the faster restatement beside it, for comparison with round 1
            pfps, pms, _, _, _ = cpu_reference_run(1, 1, args.ref_clips, kind="port")
            cpu["port"] = {"value": pfps, "unit": "frames/s", "sample": _cpu_sample_text("port", args.ref_clips, threads, cores)}

    line = {
        "metric": "frames/sec (CLIP ViT+MoCLIP+TFAM)", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(clips),
        "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e, "steps": e2e_steps},
        "gpu_launches": launches * args.steps,
        "gpu_launches_per_step": launches,
        "clocks": clocks,
        "roofline": roofline,
        "step_tflops": step_tflops,
        "step_frac_of_bf16_sustained": step_tflops / peaks["bf16_sustained"],
        "step_tflops_executed": executed_flops / (ms_step * 1e9),
        "step_flops_note": "step_tflops counts the reference's algorithmic FLOPs (all tokens of all blocks); step_tflops_executed the FLOPs of the launches "
                           "(the last block runs on the CLS rows only: same embeddings)",
        "kernel_classes": classes,
        "strong": strong,
        "ab_same_process": ab,
        "configs": other,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips", type=int, default=256, help="clips per GPU per step")
    ap.add_argument("--chunk", type=int, default=128, help="clips staged per H2D copy / tower call")
    ap.add_argument("--frames-in-flight", type=int, default=2048, help="frames per vmc_vit_forward call (workspace size)")
    ap.add_argument("--ref-clips", type=int, default=12, help="clips per CPU-baseline step (bounded sample: ~10 s per pass on 16 host threads)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling record (one 256-clip batch sharded over the N GPUs)")
    ap.add_argument("--strong-clips", type=int, default=256, help="total clips of the strong-scaling batch")
    ap.add_argument("--no-ab", action="store_true", help="skip the same-process timing of the round-1 defaults / the full last block (N = 1)")
    ap.add_argument("--no-configs", action="store_true", help="skip the short measurements of BASELINE configs 2, 3, 5")
    ap.add_argument("--ln-mode", type=int, default=0, help="tower variant (vmc_vit_model.ln_mode): 0 / 6 = bf16 residual stream + LayerNorms folded into the "
                    "qkv / c_fc GEMMs (default), 3 = fp32 stream + folds, 5 = fp32 stream + ln_1 fold, 4 = fp32 stream + separate LayerNorm kernels (round 1)")
    ap.add_argument("--last-block-cls", type=int, default=0, help="0 / 1 = the last transformer block computes only the CLS row of its output, all the "
                    "tower returns (default: identical embeddings); 2 = full last block")
    ap.add_argument("--tfam-batched", action="store_true", help="TFAM on the batched GEMM path instead of the one fused kernel")
    ap.add_argument("--prologue-impl", type=int, default=0, help="VMC_OPT_PROLOGUE_IMPL: 0 library default, 4 = 16-pixel-per-item gather kernel for uint8 frames (bit-identical output)")
    ap.add_argument("--attn-impl", type=int, default=0, help="VMC_OPT_ATTN_IMPL: 0 library default, 3 / 5 select a ViT attention kernel generation")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_main(args)
    return ours_main(args)


if __name__ == "__main__":
    sys.exit(main())
