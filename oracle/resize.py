"""numpy restatement of the resize + centre-crop the reference applies to non-224x224 frames
(TEST INFRASTRUCTURE).

Reference: ``clip_preprocess(to_pil_image(frame))`` at ``models/student_model.py:77-78`` with clip's
``_transform`` = ``Resize(224, BICUBIC)`` -> ``CenterCrop(224)``.  The arithmetic lives in Pillow
(``src/libImaging/Resample.c``, not vendored; Pillow 12.2 installed): bicubic (a = -0.5) convolution with
antialiasing support ``2 * max(scale, 1)``, coefficients computed in double, normalised to 22-bit fixed
point, a horizontal then a vertical 8-bit pass, each rounded (+2^21) and clipped to uint8.
torchvision's size rule: the short side becomes 224, the long side ``int(224 * long / short)``; the crop
offset is ``int(round((size - 224) / 2))``.  Pinned against PIL itself in ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int):
    """-> (bounds int32 [out,2] (xmin, count), coeffs int32 [out, ksize]) as Pillow's precompute_coeffs +
    normalize_coeffs_8bpc for box (0, in_size)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(img: np.ndarray, bounds: np.ndarray, kk: np.ndarray, axis: int) -> np.ndarray:
    """One 8-bit resampling pass along ``axis`` of a uint8 array."""
    src = np.moveaxis(img, axis, -1).astype(np.int64)
    out = np.empty(src.shape[:-1] + (bounds.shape[0],), dtype=np.uint8)
    for xx in range(bounds.shape[0]):
        xmin, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = (1 << (PRECISION_BITS - 1)) + (src[..., xmin:xmin + n] * kk[xx, :n].astype(np.int64)).sum(-1)
        out[..., xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, -1, axis)


def resized_size(h: int, w: int, size: int = 224):
    """torchvision ``Resize(int)`` output size (h, w)."""
    if w <= h:
        return int(size * h / w), size
    return size, int(size * w / h)


def resize_center_crop_u8(frames: np.ndarray, size: int = 224, hf_crop: bool = False) -> np.ndarray:
    """uint8 [..., H, W] -> uint8 [..., size, size]: PIL bicubic resize (short side -> size, up or down) then CenterCrop
    (torchvision: int(round(margin / 2)), Python round-half-to-even; hf_crop: HF image_transforms.center_crop, margin // 2,
    the CLIPImageProcessor of extract_embeddings.py:18,91)."""
    H, W = frames.shape[-2:]
    nh, nw = resized_size(H, W, size)
    out = frames
    if nw != W:
        out = _pass(out, *precompute_coeffs(W, nw), axis=-1)  # horizontal first (Resample.c)
    if nh != H:
        out = _pass(out, *precompute_coeffs(H, nh), axis=-2)
    top = (nh - size) // 2 if hf_crop else int(round((nh - size) / 2.0))
    left = (nw - size) // 2 if hf_crop else int(round((nw - size) / 2.0))
    return np.ascontiguousarray(out[..., top:top + size, left:left + size])
