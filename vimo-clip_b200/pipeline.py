"""Full ViMoCLIP inference chained in HBM (BASELINE config 4): CLIP ViT on RGB frames + MoCLIP student
on motion frames + TFAM fusion -> multi-label logits.  The reference decouples the three stages through
HDF5 files (extract_embeddings.py -> inference.py -> TFAM/train_and_eval.py); here the embeddings never
leave the GPU.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import sharding
from .clip_hf import CLIPVisionFeatures
from .student import FlowStudentModel, FrameDiffStudentModel
from .tfam import AMO_CLIP


class ViMoCLIPPipeline(nn.Module):
    def __init__(self, rgb_model: str = "openai/clip-vit-base-patch16", student_model: str = "ViT-B/32", num_classes: int = 140,
                 frame_diff: bool = False, device="cuda", clips_per_step: int = 128):
        super().__init__()
        self.rgb = CLIPVisionFeatures(rgb_model).to(device)
        cls = FrameDiffStudentModel if frame_diff else FlowStudentModel
        self.student = cls(student_model, device=device, num_classes=num_classes)
        self.tfam = AMO_CLIP(num_classes=num_classes, device=device).to(device).eval()
        self.clips_per_step = clips_per_step
        self.device = torch.device(device)
        self._copy_stream = None
        self.ramp_first_chunk = True  # host inputs: small first chunk (see _chunks)

    def _chunks(self, n_clips: int, on_host: bool):
        """Clip ranges of the tower calls.  Device-resident inputs: ``clips_per_step`` clips each.  Host inputs: the FIRST chunk is
        a quarter of that, because its host->device copy is the only one no kernel can overlap (128 clips of RGB frames = 308 MB,
        ~12 ms of PCIe before the first kernel starts; 32 clips = 3 ms), then full chunks."""
        cps = self.clips_per_step
        first = max(1, cps // 4) if (on_host and self.ramp_first_chunk and n_clips > cps // 4 and cps >= 4) else cps
        bounds, c0 = [], 0
        while c0 < n_clips:
            c1 = min(n_clips, c0 + (first if c0 == 0 else cps))
            bounds.append((c0, c1))
            c0 = c1
        return bounds

    def _stage(self, rgb_u8, motion_u8, c0, c1):
        """Queue the host->device copies of one clip chunk on the copy stream: RGB first, then motion, each with its own
        event, so the RGB tower starts as soon as ITS frames have arrived while the motion frames are still in flight.
        Returns (rgb, rgb_event, motion, motion_event)."""
        r = rgb_u8[c0:c1]
        m = motion_u8[c0:c1]
        if r.is_cuda and m.is_cuda:
            return r, None, m, None
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        compute = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self._copy_stream):
            r = r.to(self.device, non_blocking=True)
            ev_r = torch.cuda.Event()
            ev_r.record(self._copy_stream)
            m = m.to(self.device, non_blocking=True)
            ev_m = torch.cuda.Event()
            ev_m.record(self._copy_stream)
        # the buffers are produced on the copy stream and consumed on the compute stream
        r.record_stream(compute)
        m.record_stream(compute)
        return r, ev_r, m, ev_m

    @torch.no_grad()
    def forward(self, rgb_u8: torch.Tensor, motion_u8: torch.Tensor, mask_rgb=None, mask_flow=None):
        """rgb_u8 [N,T,3,224,224] uint8 RGB frames; motion_u8 [N,T_m,3,224,224] uint8 flow / frame-diff frames
        (host or device).  Returns (logits [N,C], rgb_emb [N,T,D], motion_emb [N,T_m,D]) on the device.

        Host inputs (ideally pinned) are copied clip-chunk by clip-chunk (``clips_per_step``) on a separate
        copy stream, one chunk ahead of the compute stream, so the H2D transfer of chunk i+1 overlaps the
        kernels of chunk i; the towers batch every frame of a chunk."""
        N, T = rgb_u8.shape[:2]
        e_rgb, e_mot = [], []
        chunks = self._chunks(N, on_host=not (rgb_u8.is_cuda and motion_u8.is_cuda))
        staged = self._stage(rgb_u8, motion_u8, *chunks[0])
        for k in range(len(chunks)):
            r, ev_r, m, ev_m = staged
            if k + 1 < len(chunks):
                staged = self._stage(rgb_u8, motion_u8, *chunks[k + 1])
            compute = torch.cuda.current_stream(self.device)
            if ev_r is not None:
                compute.wait_event(ev_r)
            n = r.shape[0]
            e_rgb.append(self.rgb.get_image_features_u8(r.reshape(n * T, *r.shape[2:])).view(n, T, -1))
            if ev_m is not None:
                compute.wait_event(ev_m)
            e_mot.append(self.student(m)[0])
        er = e_rgb[0] if len(e_rgb) == 1 else torch.cat(e_rgb)
        em = e_mot[0] if len(e_mot) == 1 else torch.cat(e_mot)
        logits = self.tfam(er, em, mask_rgb, mask_flow)
        return logits, er, em

    @torch.no_grad()
    def forward_sharded(self, rgb_u8_local, motion_u8_local, num_clips: int):
        """Each rank passes only the clips it owns (``sharding.local_clip_ids``); every rank gets the
        gathered logits / embeddings in global clip order (one NCCL all-gather each)."""
        lg, er, em = self.forward(rgb_u8_local, motion_u8_local)
        return (sharding.gather_clips(lg, num_clips), sharding.gather_clips(er, num_clips), sharding.gather_clips(em, num_clips))
