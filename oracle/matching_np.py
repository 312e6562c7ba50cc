"""Numpy restatement of the matching track's candidate-feature construction.

TEST INFRASTRUCTURE (see oracle/__init__.py): the checker for vsc22_submission_b200/matching.py +
csrc/pair_sims.cu `pair_segment_images_kernel`, never a fallback for them.

Restates (file:line under /root/reference/VSC22-Matching-Track-1st/infer):
* src/utils.py:18-47 ``generate_candidates_classfiy_feature`` and :50-73 ``generate_matching_feature``: when a query
  video holds several ``num_data``-frame copies (``num_data != len(qfeat)``), keep the copy with the largest mean of its
  10 largest row maxima of ``qfeat @ rfeat.T`` (``maxs.sort(); maxs[-10:].mean()``, ``np.argmax``: first maximum);
* src/dataset.py:103-125 ``MatchClassifyDataset.__getitem__`` (crop to / zero-pad into H x W, 3 identical channels; one
  item for ``q @ r.T`` and one for ``r @ q.T``) and :128-144 ``MatchRefineDataset.__getitem__`` (same for ``q @ r.T``
  only, plus the (h, w) actually filled).

Pinned by tests/test_oracle_matching.py against those functions / classes imported from the reference tree.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def select_segment(qfeat: np.ndarray, rfeat: np.ndarray, num_data: int) -> Tuple[int, np.ndarray]:
    """-> (index of the kept copy, its rows)."""
    if num_data == len(qfeat):
        return 0, qfeat
    sim = np.matmul(qfeat, rfeat.T)
    scores, start = [], 0
    while start < len(qfeat):
        maxs = sim[start:start + num_data].max(1)
        maxs.sort()
        scores.append(maxs[-10:].mean())
        start += num_data
    k = int(np.argmax(scores))
    return k, qfeat[k * num_data:(k + 1) * num_data]


def padded(mat: np.ndarray, resolution=(160, 160)) -> Tuple[np.ndarray, int, int]:
    h, w = min(mat.shape[0], resolution[0]), min(mat.shape[1], resolution[1])
    out = np.zeros(resolution, dtype=np.float32)
    out[:h, :w] += mat[:h, :w]
    return out, h, w


def classify_images(query: dict, ref: dict, candidate_list, query_video_len_map: dict, resolution=(160, 160)) -> np.ndarray:
    """[2 * n, H, W]: for candidate i, image 2i = q_kept @ r.T, image 2i+1 = r @ q_kept.T (each replicated to 3 channels
    by the dataset)."""
    out = []
    for qid, rid, _ in candidate_list:
        _, q = select_segment(query[qid], ref[rid], query_video_len_map[qid])
        out.append(padded(np.matmul(q, ref[rid].T), resolution)[0])
        out.append(padded(np.matmul(ref[rid], q.T), resolution)[0])
    return np.stack(out) if out else np.zeros((0, *resolution), np.float32)


def refine_images(query: dict, ref: dict, candidate_list, query_video_len_map: dict, resolution=(224, 224)):
    """([n, H, W], segment [n], h [n], w [n])."""
    imgs, segs, hs, ws = [], [], [], []
    for qid, rid, _ in candidate_list:
        k, q = select_segment(query[qid], ref[rid], query_video_len_map[qid])
        im, h, w = padded(np.matmul(q, ref[rid].T), resolution)
        imgs.append(im); segs.append(k); hs.append(h); ws.append(w)
    return (np.stack(imgs) if imgs else np.zeros((0, *resolution), np.float32)), np.array(segs), np.array(hs), np.array(ws)
