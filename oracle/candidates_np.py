"""Numpy restatement of the reference's video-level candidate generation.

TEST INFRASTRUCTURE (see oracle/__init__.py): imported only by tests/, __graft_entry__.smoke() and the CPU-baseline
legs of bench.py.  It is the checker for vsc22_submission_b200/candidates.py + csrc/global_topk.cu, never a
fallback for them.

Restates (file:line under /root/reference):
* VSC22-Descriptor-Track-1st/infer/vsc/index.py:142-165 ``VideoIndex._global_threshold_knn_search`` over
  vsc/exhaustive_search.py:206-292 ``range_search_max_results``: an adaptive-radius range search that keeps between
  ``global_k`` and ``2*global_k`` of the best (query frame, ref frame) pairs seen so far, then sorts them by score
  (stable, so ties stay in (query row, result order) order) and keeps the first ``global_k``.  Every pair better than
  the final radius survives the radius updates, so the outcome is the exact global top-K of all frame pairs
  (``global_topk_pairs`` below computes it directly); the two differ only when scores tie exactly AT the radius, where
  the reference's strict comparison (exhaustive_search.py:151-155) drops all tied pairs.
* vsc/index.py:119-140: hits regrouped per (query video, ref video) in order of first appearance.
* vsc/candidates.py:24-40: ``MaxScoreAggregation`` + ``sorted(..., reverse=True)`` (stable).
* VSC22-Matching-Track-1st/infer/infer_matching.py:229-256: per-video search, hits above ``SEARCH_THRESHOLD``,
  dict max-reduce, ``sort(key=-score)``.

Pinned by tests/test_oracle_candidates.py against the reference's own ``CandidateGeneration`` (run in-container over the
faiss stand-in) and against ``tests/golden/search_small.npz`` (cand_q / cand_r / cand_s written by the reference).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

from . import faiss_np


def _scores(queries: np.ndarray, refs: np.ndarray, metric: int) -> np.ndarray:
    ix = faiss_np.IndexFlat(refs.shape[1], metric)
    ix.add(refs)
    return ix._scores(np.ascontiguousarray(queries, dtype=np.float32))


def global_topk_pairs(queries: np.ndarray, refs: np.ndarray, global_k: int,
                      metric: int = faiss_np.METRIC_INNER_PRODUCT) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(score, query row, ref row) of the ``global_k`` best frame pairs, best first, ties by (query row, ref row)."""
    S = _scores(queries, refs, metric)
    flat = S.reshape(-1)
    key = -flat if metric == faiss_np.METRIC_INNER_PRODUCT else flat
    order = np.argsort(key, kind="stable")[: max(int(global_k), 0)]
    return flat[order], order // S.shape[1], order % S.shape[1]


def threshold_pairs(queries: np.ndarray, refs: np.ndarray, threshold: float,
                    metric: int = faiss_np.METRIC_INNER_PRODUCT) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """All frame pairs strictly better than ``threshold`` (faiss range_search semantics), best first."""
    S = _scores(queries, refs, metric)
    flat = S.reshape(-1)
    keep_max = metric == faiss_np.METRIC_INNER_PRODUCT
    hit = np.flatnonzero(flat > threshold if keep_max else flat < threshold)
    order = hit[np.argsort(-flat[hit] if keep_max else flat[hit], kind="stable")]
    return flat[order], order // S.shape[1], order % S.shape[1]


def _offsets(lengths: Sequence[int]) -> np.ndarray:
    off = np.zeros(len(lengths) + 1, dtype=np.int64)
    np.cumsum(np.asarray(lengths, dtype=np.int64), out=off[1:])
    return off


def video_pair_candidates(scores: np.ndarray, qrows: np.ndarray, rrows: np.ndarray, q_len: Sequence[int],
                          r_len: Sequence[int]) -> List[Tuple[int, int, float]]:
    """(query video, ref video, best frame-pair score), sorted by score descending with ties in order of first
    appearance -- index.py:119-140 + candidates.py:36-40 with MaxScoreAggregation."""
    q_off, r_off = _offsets(q_len), _offsets(r_len)
    qv = np.searchsorted(q_off, qrows, side="right") - 1
    rv = np.searchsorted(r_off, rrows, side="right") - 1
    best: dict = {}
    for s, a, b in zip(scores, qv, rv):                    # dict keeps first-appearance order
        k = (int(a), int(b))
        if k not in best or s > best[k]:
            best[k] = s
    cands = [(a, b, s) for (a, b), s in best.items()]
    return sorted(cands, key=lambda c: c[2], reverse=True)


def candidates(queries: np.ndarray, refs: np.ndarray, q_len: Sequence[int], r_len: Sequence[int], global_k: int,
               metric: int = faiss_np.METRIC_INNER_PRODUCT) -> List[Tuple[int, int, float]]:
    """``CandidateGeneration(refs, MaxScoreAggregation()).query(queries, global_k)`` on row-concatenated arrays."""
    s, qi, ri = global_topk_pairs(queries, refs, global_k, metric)
    if metric != faiss_np.METRIC_INNER_PRODUCT:
        raise NotImplementedError("candidates.py sorts by descending score; the reference only uses it with IP")
    return video_pair_candidates(s, qi, ri, q_len, r_len)


def threshold_candidates(queries: np.ndarray, refs: np.ndarray, q_len: Sequence[int], r_len: Sequence[int],
                         threshold: float) -> List[Tuple[int, int, float]]:
    """``search_res_list`` of infer_matching.py:229-256."""
    s, qi, ri = threshold_pairs(queries, refs, threshold)
    return video_pair_candidates(s, qi, ri, q_len, r_len)
