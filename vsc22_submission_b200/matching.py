"""Matching-track candidate features on the device.

Mirrors ``generate_candidates_classfiy_feature`` / ``generate_matching_feature``
(VSC22-Matching-Track-1st/infer/src/utils.py:18-47, 50-73) fused with the batching of ``MatchClassifyDataset`` /
``MatchRefineDataset`` (infer/src/dataset.py:103-144): for every (query_id, ref_id, score) candidate the frame-similarity
matrix of the best query copy, cropped / zero-padded to the CNN's input resolution, leaves the GPU-resident descriptor
arrays as ONE ``[N, 3, H, W]`` CUDA batch (the 3 channels are an expanded view) -- what
``match_classify`` / ``match_refine`` (infer_matching.py:158-204) feed to their models.  The reference does a numpy
matmul per candidate twice (selection, then features), keeps every matrix in host lists and re-uploads them through a
DataLoader.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .localization import PairSimilarity, _p


@dataclasses.dataclass
class _Video:
    video_id: str
    feature: np.ndarray


@dataclasses.dataclass
class _Pair:
    query_id: str
    ref_id: str


class MatchingFeatures(PairSimilarity):
    """``MatchingFeatures(query_map, ref_map)`` with ``{video_id: feature [n, d]}`` dicts (infer_matching.py:268-269)."""

    def __init__(self, query: Dict[str, np.ndarray], ref: Dict[str, np.ndarray], device="cuda"):
        super().__init__([_Video(k, v) for k, v in query.items()], [_Video(k, v) for k, v in ref.items()], 0.0, device)

    def _images(self, candidate_list: Sequence, query_video_len_map: Dict[str, int], resolution, with_transpose: bool):
        H, W = int(resolution[0]), int(resolution[1])
        n, n_img = len(candidate_list), 2 if with_transpose else 1
        dev = self.device
        if n == 0:
            return torch.zeros((0, n_img, H, W), device=dev), np.zeros((0, 4), np.int32)
        pairs = [_Pair(c[0], c[1]) for c in candidate_list]
        sims, s_off, row_off, ql, rl, _, _ = self._device_run(pairs, 0)
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
        seg = t([query_video_len_map[c[0]] for c in candidate_list], np.int32)
        d_ql, d_rl, d_so, d_row = t(ql, np.int32), t(rl, np.int32), t(s_off[:-1], np.int64), t(row_off[:-1], np.int64)
        rowmax = torch.empty((int(row_off[-1]),), dtype=torch.float32, device=dev)
        images = torch.empty((n, n_img, H, W), dtype=torch.float32, device=dev)
        info = torch.empty((n, 4), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(_lib.lib().vscb200_pair_segment_images(_p(sims), n, _p(d_ql), _p(d_rl), _p(d_so), _p(d_row), _p(seg),
                                                              _p(rowmax), H, W, 1 if with_transpose else 0, _p(images),
                                                              _p(info), stream), "pair_segment_images")
        return images, info.cpu().numpy()

    def classify_batch(self, candidate_list: Sequence, query_video_len_map: Dict[str, int], resolution=(160, 160)
                       ) -> Tuple[torch.Tensor, List[list]]:
        """-> (features [2n, 3, H, W] CUDA, infos) in the order of utils.py:43-46: for candidate i, item 2i is
        ``q @ r.T`` and item 2i+1 ``r @ q.T``; ``infos[j] = [qid, rid, score]``."""
        images, _ = self._images(candidate_list, query_video_len_map, resolution, True)
        n = images.shape[0]
        feats = images.reshape(2 * n, 1, *images.shape[2:]).expand(-1, 3, -1, -1)
        infos = [[c[0], c[1], c[2]] for c in candidate_list for _ in (0, 1)]
        return feats, infos

    def refine_batch(self, candidate_score_list: Sequence, query_video_len_map: Dict[str, int], resolution=(224, 224)):
        """-> (features [n, 3, H, W] CUDA, qids, rids, h [n], w [n], kept copy [n]) -- the batches of
        ``MatchRefineDataset`` (dataset.py:128-144) over ``generate_matching_feature``'s list."""
        images, info = self._images(candidate_score_list, query_video_len_map, resolution, False)
        feats = images.expand(-1, 3, -1, -1)
        return (feats, [c[0] for c in candidate_score_list], [c[1] for c in candidate_score_list], info[:, 2].copy(),
                info[:, 3].copy(), info[:, 0].copy())
