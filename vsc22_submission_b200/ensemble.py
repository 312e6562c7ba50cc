"""Ensemble tail between the encoders and the index, device-resident.

Mirrors what the reference does on the host with sklearn, one video at a time:

* ``sklearn.preprocessing.normalize(x)`` per model, ``np.concatenate(..., axis=1)``, ``pca_model.transform(x)``
  -- D/infer/concat_pca_sn.py:56-64, D/infer/extract_query_feats.py:169-204, M/infer/infer_matching.py:140-145.

``B200PCA`` is constructed from the two arrays a fitted ``sklearn.decomposition.PCA`` carries (``mean_``,
``components_``; the reference pickles the sklearn object, concat_pca_sn.py:47-51) and offers the same
``transform(X)`` call for host arrays, plus ``transform_parts`` which fuses the per-model normalisation and the
concatenation for descriptors that are still on the device.  All arithmetic is fp32 in libvscb200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np
import torch

from . import _lib


class B200PCA:
    def __init__(self, mean, components, device="cuda"):
        self.device = torch.device(device)
        self.mean_ = torch.as_tensor(np.asarray(mean, dtype=np.float32)).to(self.device).contiguous()
        self.components_ = torch.as_tensor(np.asarray(components, dtype=np.float32)).to(self.device).contiguous()
        if self.components_.dim() != 2 or self.mean_.numel() != self.components_.shape[1]:
            raise ValueError("B200PCA: components_ must be [n_components, n_features] and mean_ [n_features]")
        self.n_components_, self.n_features_ = int(self.components_.shape[0]), int(self.components_.shape[1])

    @classmethod
    def from_sklearn(cls, pca, device="cuda"):
        if getattr(pca, "whiten", False):
            raise ValueError("B200PCA: whitened PCA models are not on the reference's path (PCA(n_components=512))")
        return cls(pca.mean_, pca.components_, device)

    def _run(self, parts: Sequence[torch.Tensor]) -> torch.Tensor:
        n = int(parts[0].shape[0])
        dims = [int(p.shape[1]) for p in parts]
        if sum(dims) != self.n_features_ or any(int(p.shape[0]) != n for p in parts):
            raise AssertionError(f"expected {self.n_features_} features in total and equal row counts, got {dims}")
        parts = [p.contiguous().float() for p in parts]
        out = torch.empty((n, self.n_components_), dtype=torch.float32, device=self.device)
        if n == 0:
            return out
        ptrs = (C.c_void_p * len(parts))(*[p.data_ptr() for p in parts])
        dims_c = (C.c_int * len(parts))(*dims)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(_lib.lib().vscb200_ensemble_pca(ptrs, dims_c, len(parts), n, C.c_void_p(self.mean_.data_ptr()),
                                                       C.c_void_p(self.components_.data_ptr()), self.n_components_,
                                                       C.c_void_p(out.data_ptr()), C.c_void_p(stream)), "ensemble_pca")
        return out

    def transform_parts(self, parts: Sequence[torch.Tensor]) -> torch.Tensor:
        """normalize(part) for every model's descriptors [n, d_i] (CUDA tensors) -> concat -> PCA.transform."""
        for p in parts:
            if not isinstance(p, torch.Tensor) or not p.is_cuda:
                raise RuntimeError("B200PCA.transform_parts: expected CUDA tensors (no CPU fallback)")
        return self._run(list(parts))

    def transform_parts_host(self, parts: Sequence[np.ndarray]) -> np.ndarray:
        dev = [torch.from_numpy(np.ascontiguousarray(p, dtype=np.float32)).to(self.device) for p in parts]
        return self._run(dev).cpu().numpy()


def near_dup_keep(features: torch.Tensor, threshold: float = 0.975) -> torch.Tensor:
    """Near-duplicate frame filter of one query video (D/infer/extract_query_feats.py:190-200, ``FRAME_THRESHOLD``):
    bool CUDA mask [n] of the frames that survive -- ``to_keep_idx`` of the reference as a mask."""
    if not isinstance(features, torch.Tensor) or not features.is_cuda:
        raise RuntimeError("near_dup_keep: expected a CUDA tensor (no CPU fallback)")
    x = features.contiguous().float()
    keep = torch.empty((x.shape[0],), dtype=torch.uint8, device=x.device)
    if x.shape[0]:
        with torch.cuda.device(x.device):
            stream = torch.cuda.current_stream(x.device).cuda_stream
            _lib.check(_lib.lib().vscb200_near_dup_keep(C.c_void_p(x.data_ptr()), x.shape[0], x.shape[1], float(threshold),
                                                        C.c_void_p(keep.data_ptr()), C.c_void_p(stream)), "near_dup_keep")
    return keep.bool()


def query_tail(pca: B200PCA, parts: Sequence[torch.Tensor], frame_threshold: float = 0.975):
    """The descriptor tail of ``Main.process`` for a query video that passed the video-score gate
    (extract_query_feats.py:169-204): per-model normalize -> concat -> drop near-duplicate frames -> PCA.
    Returns (features [n_kept, n_components] CUDA, kept frame indices int64 CUDA)."""
    normed = [torch.nn.functional.normalize(p.float(), dim=1) for p in parts]         # sklearn normalize per model
    keep = near_dup_keep(torch.cat(normed, dim=1), frame_threshold)
    idx = torch.nonzero(keep).flatten()
    return pca.transform_parts([p[idx] for p in parts]), idx
