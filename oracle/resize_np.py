"""Numpy restatement of the frame preprocessing the reference applies before its encoders.

TEST INFRASTRUCTURE (see oracle/__init__.py): the checker for vsc22_submission_b200/ingest.py + csrc/resize.cu.

Reference call sites: VSC22-Descriptor-Track-1st/infer/src/transform.py:20-43 (``sscd_transform`` / ``eff_transform`` /
``vit_transform``: ``Resize([w, h], interpolation=BICUBIC)`` -> ``ToTensor()`` -> ``Normalize(mean, std)`` applied to the
PIL image of every decoded frame, infer/src/dataset.py:126-155; the same Compose objects at
extract_query_feats.py:96-125).  The arithmetic lives in third-party code that is not in the reference tree:

* Pillow (version unpinned in the reference's dockerfile; 12.2.0 in this image) ``Image.resize(..., BICUBIC)`` =
  libImaging/Resample.c ``ImagingResample``: separable, antialiased (filter support scaled by the down-scale factor),
  horizontal pass then vertical pass, 8-bit intermediates.  Restated here from its published algorithm:
  bicubic kernel with a = -0.5 and support 2; per output coordinate ``center = (xx + 0.5) * scale``, taps
  ``[int(center - support + 0.5), int(center + support + 0.5))`` clamped to the image, weights
  ``filter((x - center + 0.5) / filterscale)`` normalised to sum 1 in double precision, quantised to 22-bit fixed point
  (``int(+-0.5 + w * 2**22)``); a pixel = ``clip8((2**21 + sum(tap * w)) >> 22)``.
* torchvision ``ToTensor`` (uint8 HWC -> float32 CHW / 255) and ``Normalize`` ((x - mean) / std in float32).

Pinned by tests/test_oracle_resize.py against Pillow + torchvision through the reference's own ``sscd_transform`` /
``vit_transform`` (imported from the reference tree) and against tests/golden/resize_small.npz written by them.
"""
from __future__ import annotations

import math
from typing import Sequence, Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def coefficients(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """(bounds int32 [out, 2] = (first tap, tap count), weights int32 [out, ksize], ksize) -- precompute_coeffs +
    normalize_coeffs_8bpc of Resample.c for the full-image box."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def _pass(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """One resampling pass of a uint8 [H, W, C] image along `axis` (0 = vertical, 1 = horizontal)."""
    bounds, kk, _ = coefficients(img.shape[axis], out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], np.uint8)
    for xx in range(out_size):
        x0, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(kk[xx, :n].astype(np.int64), src[x0:x0 + n], axes=(0, 0))
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bicubic_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """``PIL.Image.fromarray(img).resize((out_w, out_h), BICUBIC)`` for a uint8 [H, W, C] array."""
    if img.shape[1] != out_w:
        img = _pass(img, out_w, 1)
    if img.shape[0] != out_h:
        img = _pass(img, out_h, 0)
    return img


def to_tensor_normalize(img_u8: np.ndarray, mean: Sequence[float], std: Sequence[float]) -> np.ndarray:
    """``Normalize(mean, std)(ToTensor()(img))`` -> float32 [C, H, W]."""
    x = img_u8.transpose(2, 0, 1).astype(np.float32) / np.float32(255)
    m = np.asarray(mean, np.float32)[:, None, None]
    s = np.asarray(std, np.float32)[:, None, None]
    return (x - m) / s


def preprocess(img_u8: np.ndarray, out_h: int, out_w: int, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)) -> np.ndarray:
    return to_tensor_normalize(resize_bicubic_u8(img_u8, out_h, out_w), mean, std)
