"""numpy restatement of the integer / fp32 prologue (TEST INFRASTRUCTURE).

* ``to_pil_u8``          -- ``torchvision.transforms.functional.to_pil_image`` on a float CHW tensor:
                            ``pic.mul(255).byte()``; float->uint8 goes through int64 truncation and
                            keeps the low 8 bits (models/student_model.py:74,78).
* ``normalise_u8``       -- ``ToTensor`` (``float32(u8)/255``) + ``Normalize`` (``(v-mean)/std``), true
                            fp32 divisions (clip ``_transform``).
* ``preprocess_frames``  -- the whole in-forward preprocessing at 224x224 for the three input regimes.
* ``bgr2gray`` / ``frame_difference`` -- ``cv2.cvtColor(BGR2GRAY)`` + ``cv2.absdiff``
                            (utils/generate_frame_diff_video.py:37,46,49) in closed integer form.
* ``patchify``           -- [F,3,H,W] -> [F*n, 3*p*p] in ``conv1.weight.reshape(d, -1)`` column order.
"""
from __future__ import annotations

import numpy as np

CLIP_MEAN = np.array([0.48145466, 0.4578275, 0.40821073], dtype=np.float32)
CLIP_STD = np.array([0.26862954, 0.26130258, 0.27577711], dtype=np.float32)


def to_pil_u8(frames: np.ndarray) -> np.ndarray:
    """float or uint8 [..] -> uint8 as ``frame.float()`` + ``to_pil_image`` would produce."""
    x = frames.astype(np.float32)  # student_model.py:74 .float()
    y = x * np.float32(255.0)  # to_pil_image: pic.mul(255)
    return (np.trunc(y).astype(np.int64) & 255).astype(np.uint8)  # .byte(): via int64, low 8 bits


def normalise_u8(u8: np.ndarray) -> np.ndarray:
    """uint8 [F,3,H,W] -> fp32 [F,3,H,W] = (u8/255 - mean)/std with fp32 true division."""
    t = u8.astype(np.float32) / np.float32(255.0)
    return ((t - CLIP_MEAN[None, :, None, None]) / CLIP_STD[None, :, None, None]).astype(np.float32)


def preprocess_frames(frames: np.ndarray) -> np.ndarray:
    """[F,3,H,W] uint8 (regime A) or float (regimes B, C) -> normalised fp32 [F,3,224,224] fed to the ViT:
    to_pil_image wrap -> Resize(224, bicubic) -> CenterCrop(224) (identities at 224x224) -> ToTensor -> Normalize."""
    u8 = to_pil_u8(frames)
    if u8.shape[-1] != 224 or u8.shape[-2] != 224:
        from .resize import resize_center_crop_u8

        u8 = resize_center_crop_u8(u8, 224)
    return normalise_u8(u8)


def bgr2gray(bgr: np.ndarray) -> np.ndarray:
    """OpenCV 8-bit BGR2GRAY: (B*3735 + G*19235 + R*9798 + 2^14) >> 15, [...,3] uint8 -> [...] uint8."""
    b = bgr[..., 0].astype(np.int64)
    g = bgr[..., 1].astype(np.int64)
    r = bgr[..., 2].astype(np.int64)
    return ((b * 3735 + g * 19235 + r * 9798 + 16384) >> 15).astype(np.uint8)


def frame_difference(bgr_frames: np.ndarray) -> np.ndarray:
    """[T+1,H,W,3] uint8 BGR -> [T,H,W] uint8 = |gray[t+1] - gray[t]| (generate_frame_diff_video.py:37-49)."""
    gray = bgr2gray(bgr_frames).astype(np.int16)
    return np.abs(gray[1:] - gray[:-1]).astype(np.uint8)


def patchify(x: np.ndarray, patch: int) -> np.ndarray:
    """[F,3,H,W] -> [F*(H/p)*(W/p), 3*p*p]; column = c*p*p + iy*p + ix (== conv1.weight.reshape(d,-1))."""
    F, C, H, W = x.shape
    gh, gw = H // patch, W // patch
    y = x.reshape(F, C, gh, patch, gw, patch).transpose(0, 2, 4, 1, 3, 5)
    return np.ascontiguousarray(y).reshape(F * gh * gw, C * patch * patch)
