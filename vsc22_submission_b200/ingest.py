"""Frame preprocessing on the device (the part after JPEG decode).

Mirrors the transform factories of VSC22-Descriptor-Track-1st/infer/src/transform.py:20-43 (``sscd_transform``,
``eff_transform``, ``vit_transform``: ``Resize([width, height], BICUBIC)`` -> ``ToTensor()`` -> ``Normalize``), which the
reference applies frame by frame on PIL images inside CPU DataLoader workers (infer/src/dataset.py:126-155,
extract_query_feats.py:96-125).  Here the decoded uint8 frames of a video are uploaded once (1 byte per sample instead of
the 4-byte float tensor the reference ships) and resized + normalised by csrc/resize.cu with Pillow's exact fixed-point
arithmetic; the result is the ``[n, 3, h, w]`` float32 CUDA tensor the encoders take.

The decode half runs on the device too: ``decode_jpeg_frames`` takes the JPEG files of a video's frames (the ``bytes`` the
reference reads from the zip, dataset.py:137-139) and returns the RGB frames as a CUDA tensor, bit-identical to
``PIL.Image.open(io.BytesIO(b)).convert("RGB")`` (csrc/jpeg.cu); ``FramePreprocessor`` accepts those ``bytes`` directly and
``video_zip_frames`` mirrors ``D_vsc.__getitem__`` (dataset.py:126-148) without a PIL image in between.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence, Union

import numpy as np
import torch

from . import _lib

IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def decode_jpeg_frames(files: Sequence[bytes], device="cuda") -> torch.Tensor:
    """The JPEG files of equally sized frames -> uint8 CUDA tensor ``[n, H, W, 3]`` (RGB), bit-identical to Pillow's decode.
    Baseline YCbCr (4:2:0 / 4:2:2 / 4:4:4) or grey files, as ffmpeg writes them; anything else raises (no CPU fallback)."""
    device = torch.device(device)
    n = len(files)
    if n == 0:
        return torch.empty((0, 0, 0, 3), dtype=torch.uint8, device=device)
    bufs = [bytes(f) if not isinstance(f, (bytes, bytearray)) else f for f in files]
    ptrs = (C.c_char_p * n)(*bufs)                 # no copy: ctypes passes the bytes objects' own buffers
    sizes = (C.c_uint64 * n)(*[len(b) for b in bufs])
    h, w = C.c_int(0), C.c_int(0)
    lib = _lib.lib()
    with torch.cuda.device(device):
        stream = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        _lib.check(lib.vscb200_jpeg_decode(ptrs, sizes, 1, None, C.byref(h), C.byref(w), stream), "jpeg_decode (header)")
        out = torch.empty((n, h.value, w.value, 3), dtype=torch.uint8, device=device)
        _lib.check(lib.vscb200_jpeg_decode(ptrs, sizes, n, C.c_void_p(out.data_ptr()), C.byref(h), C.byref(w), stream), "jpeg_decode")
    return out


def video_zip_frames(zip_path: str, preprocess: "FramePreprocessor") -> torch.Tensor:
    """``D_vsc.__getitem__`` (VSC22-Descriptor-Track-1st/infer/src/dataset.py:126-148) on the device: the frames of one video
    -- the JPEG members of its zip in sorted name order -- decoded, resized and normalised -> ``[n, 3, h, w]`` float32."""
    from zipfile import ZipFile
    with ZipFile(zip_path, "r") as z:
        files = [z.read(name) for name in sorted(z.namelist())]
    return preprocess(files)


class FramePreprocessor:
    """Callable on ``[n, H, W, 3]`` uint8 frames (numpy, CUDA tensor, or a sequence of equally sized PIL images / arrays);
    returns float32 CUDA ``[n, 3, width, height]`` -- torchvision's ``Resize([width, height])`` reads its argument as
    (rows, columns), so does this."""

    def __init__(self, width: int, height: int, mean: Sequence[float], std: Sequence[float], device="cuda"):
        self.out_h, self.out_w = int(width), int(height)
        self.mean = (C.c_float * 3)(*[float(m) for m in mean])
        self.std = (C.c_float * 3)(*[float(s) for s in std])
        self.device = torch.device(device)

    def _frames(self, frames) -> torch.Tensor:
        if isinstance(frames, torch.Tensor):
            t = frames
        elif isinstance(frames, (list, tuple)) and len(frames) and isinstance(frames[0], (bytes, bytearray)):
            return decode_jpeg_frames(frames, self.device)          # JPEG files: decoded on the device
        else:
            if not isinstance(frames, np.ndarray):
                frames = np.stack([np.asarray(f) for f in frames])
            t = torch.from_numpy(np.ascontiguousarray(frames))
        if t.dim() == 3:
            t = t[None]
        if t.dtype != torch.uint8 or t.dim() != 4 or t.shape[-1] != 3:
            raise AssertionError(f"expected uint8 frames [n, H, W, 3], got {t.dtype} {tuple(t.shape)}")
        return t.to(self.device, non_blocking=True).contiguous()

    def __call__(self, frames: Union[np.ndarray, torch.Tensor, Sequence]) -> torch.Tensor:
        x = self._frames(frames)
        n, H, W = int(x.shape[0]), int(x.shape[1]), int(x.shape[2])
        out = torch.empty((n, 3, self.out_h, self.out_w), dtype=torch.float32, device=self.device)
        mid = torch.empty((n, H, self.out_w, 3), dtype=torch.uint8, device=self.device) if W != self.out_w else None
        with torch.cuda.device(self.device):
            stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(_lib.lib().vscb200_resize_normalize(
                C.c_void_p(x.data_ptr()), n, H, W, self.out_h, self.out_w, self.mean, self.std,
                C.c_void_p(mid.data_ptr()) if mid is not None else None, C.c_void_p(out.data_ptr()), stream),
                "resize_normalize")
        return out


def sscd_transform(width: int, height: int, device="cuda") -> FramePreprocessor:     # transform.py:20-30
    return FramePreprocessor(width, height, IMAGENET_MEAN, IMAGENET_STD, device)


def eff_transform(width: int, height: int, device="cuda") -> FramePreprocessor:      # transform.py:31-36
    return FramePreprocessor(width, height, (0.5, 0.5, 0.5), (0.5, 0.5, 0.5), device)


def vit_transform(width: int, height: int, device="cuda") -> FramePreprocessor:      # transform.py:37-43
    return FramePreprocessor(width, height, (0.5, 0.5, 0.5), (0.5, 0.5, 0.5), device)
