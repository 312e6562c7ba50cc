"""Frame-similarity matrices of all candidate video pairs on the device.

Mirrors ``LocalizationWithMetadata.similarity`` / ``VCSLLocalization.similarity`` and the list built in
``VCSLLocalization.localize_all`` (vsc/baseline/localization.py:27-66): for every ``CandidatePair`` the matrix
``queries[c.query_id].feature @ refs[c.ref_id].feature.T + similarity_bias``; in addition the per-row top-k that
``vcsl.vta.tn`` computes first (vta.py:262-265).  The reference does one numpy matmul per pair inside a 16-process
pool; here every video's descriptors are uploaded once and all pairs are computed by one launch.

``VCSLLocalization`` / ``VCSLLocalizationMaxSim`` / ``VCSLLocalizationCandidateScore`` (localization.py:38-95, as
constructed by sscd_baseline.py:113-131 with ``model_type="TN"``) run the whole of ``localize_all`` on the device: the
similarity matrices, their per-row top-k, the temporal-network alignment ``vcsl.vta.tn`` (vta.py:244-363; one warp per
pair, csrc/tn_align.cu) and the MaxSim box scores; the host only turns the boxes into ``Match`` tuples.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, NamedTuple, Sequence, Tuple

import numpy as np
import torch

from . import _lib


def _p(t):
    return C.c_void_p(t.data_ptr())


class PairSimilarity:
    """``PairSimilarity(queries, refs).similarities(candidates)`` -> [(key, sims ndarray)] like localize_all builds."""

    def __init__(self, queries: Sequence, refs: Sequence, similarity_bias: float = 0.0, device="cuda"):
        self.device = torch.device(device)
        self.bias = float(similarity_bias)
        self.q_index, self.q_feat = self._pack(queries)
        self.r_index, self.r_feat = self._pack(refs)
        if self.q_feat.shape[1] != self.r_feat.shape[1]:
            raise AssertionError("query and reference descriptors differ in dimension")

    def _pack(self, videos) -> Tuple[Dict[str, Tuple[int, int]], torch.Tensor]:
        index, off = {}, 0
        for v in videos:
            n = int(v.feature.shape[0])
            index[v.video_id] = (off, n)
            off += n
        arr = np.concatenate([np.asarray(v.feature, dtype=np.float32) for v in videos], axis=0)
        return index, torch.from_numpy(np.ascontiguousarray(arr)).to(self.device)

    def _device_run(self, candidates, top_k: int):
        qo, ql, ro, rl = [], [], [], []
        for c in candidates:
            a, b = self.q_index[c.query_id], self.r_index[c.ref_id]
            qo.append(a[0]); ql.append(a[1]); ro.append(b[0]); rl.append(b[1])
        ql_np, rl_np = np.asarray(ql, np.int64), np.asarray(rl, np.int64)
        s_off = np.concatenate([[0], np.cumsum(ql_np * rl_np)]).astype(np.int64)
        row_off = np.concatenate([[0], np.cumsum(ql_np)]).astype(np.int64)
        dev = self.device
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
        d_qo, d_ql, d_ro, d_rl = t(qo, np.int64), t(ql, np.int32), t(ro, np.int64), t(rl, np.int32)
        d_so, d_row = t(s_off[:-1], np.int64), t(row_off[:-1], np.int64)
        sims = torch.empty((int(s_off[-1]),), dtype=torch.float32, device=dev)
        n = len(qo)
        with torch.cuda.device(dev):
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(_lib.lib().vscb200_pair_sims(_p(self.q_feat), _p(self.r_feat), int(self.q_feat.shape[1]), n, _p(d_qo),
                                                    _p(d_ql), _p(d_ro), _p(d_rl), _p(d_so), self.bias, _p(sims), stream),
                       "pair_sims")
            topv = topi = None
            if top_k:
                topv = torch.empty((int(row_off[-1]), top_k), dtype=torch.float32, device=dev)
                topi = torch.empty((int(row_off[-1]), top_k), dtype=torch.int32, device=dev)
                _lib.check(_lib.lib().vscb200_pair_topk(_p(sims), n, _p(d_ql), _p(d_rl), _p(d_so), _p(d_row), int(top_k),
                                                        _p(topv), _p(topi), stream), "pair_topk")
        return sims, s_off, row_off, ql_np, rl_np, topv, topi

    def similarities(self, candidates: Sequence) -> List[Tuple[str, np.ndarray]]:
        if not candidates:
            return []
        sims, s_off, _, ql, rl, _, _ = self._device_run(candidates, 0)
        host = sims.cpu().numpy()
        return [(f"{c.query_id}-{c.ref_id}", host[s_off[i]:s_off[i + 1]].reshape(ql[i], rl[i])) for i, c in enumerate(candidates)]

    def similarities_topk(self, candidates: Sequence, top_k: int = 5):
        """-> [(key, sims, topk_indices [q_len, min(k, r_len)], topk_sims)] (vta.py:262-265)."""
        if not candidates:
            return []
        sims, s_off, row_off, ql, rl, topv, topi = self._device_run(candidates, top_k)
        host, hv, hi = sims.cpu().numpy(), topv.cpu().numpy(), topi.cpu().numpy()
        out = []
        for i, c in enumerate(candidates):
            top = min(top_k, int(rl[i]))
            rows = slice(int(row_off[i]), int(row_off[i + 1]))
            out.append((f"{c.query_id}-{c.ref_id}", host[s_off[i]:s_off[i + 1]].reshape(ql[i], rl[i]),
                        hi[rows, :top].astype(np.int64), hv[rows, :top]))
        return out


class _Match(NamedTuple):                # vsc/metrics.py:182-191
    query_id: str
    ref_id: str
    score: float
    query_start: float
    query_end: float
    ref_start: float
    ref_end: float


def _match_type():
    try:
        from vsc.metrics import Match
        return Match
    except Exception:
        return _Match


def _timestamps(ts: np.ndarray, idx: int):
    """VideoMetadata.get_timestamps (vsc/index.py:26-30)."""
    t = ts[idx]
    if ts.ndim == 1:
        return (t, t)
    return (t[0], t[1])


class VCSLLocalization(PairSimilarity):
    """``VCSLLocalization(queries, refs, model_type, similarity_bias=0.0, **kwargs)`` (localization.py:38-84) for
    ``model_type == "TN"``; ``kwargs`` are ``TnVtaModel``'s (vta.py:499-518): tn_max_step, tn_top_k, max_path, min_sim,
    min_length, max_iou (``concurrency`` is accepted and ignored: there is no process pool)."""

    def __init__(self, queries: Sequence, refs: Sequence, model_type: str = "TN", similarity_bias: float = 0.0,
                 concurrency: int = 4, version: str = "v1", tn_max_step: int = 10, tn_top_k: int = 5, max_path: int = 10,
                 min_sim: float = 0.2, min_length: float = 5, max_iou: float = 0.3, device="cuda"):
        if model_type != "TN":
            raise ValueError(f"VCSLLocalization: only the temporal network ('TN') runs on the device, got {model_type!r}")
        super().__init__(queries, refs, similarity_bias, device)
        self.similarity_bias = float(similarity_bias)
        self.queries = {m.video_id: m for m in queries}
        self.refs = {m.video_id: m for m in refs}
        self.tn = dict(tn_max_step=int(tn_max_step), tn_top_k=int(tn_top_k), max_path=int(max_path),
                       min_sim=float(min_sim), min_length=float(min_length), max_iou=float(max_iou))

    def align(self, candidates: Sequence):
        """-> (boxes int32 [n, max_path + 1, 4], n_boxes int32 [n], MaxSim box scores f32 [n, max_path + 1]) numpy;
        ``boxes[i, :n_boxes[i]]`` is what ``self.model.forward_sim(sims)[i][1]`` holds in the reference."""
        cap = self.tn["max_path"] + 1
        if not candidates:
            return np.zeros((0, cap, 4), np.int32), np.zeros((0,), np.int32), np.zeros((0, cap), np.float32)
        k = self.tn["tn_top_k"]
        sims, s_off, row_off, ql, rl, topv, topi = self._device_run(candidates, k)
        dev, n = self.device, len(candidates)
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
        d_ql, d_rl, d_row, d_so = t(ql, np.int32), t(rl, np.int32), t(row_off[:-1], np.int64), t(s_off[:-1], np.int64)
        boxes = torch.zeros((n, cap, 4), dtype=torch.int32, device=dev)
        nb = torch.zeros((n,), dtype=torch.int32, device=dev)
        score = torch.zeros((n, cap), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(_lib.lib().vscb200_tn_align(_p(topv), _p(topi), k, n, _p(d_ql), _p(d_rl), _p(d_row), int(ql.max()),
                                                   self.tn["tn_max_step"], self.tn["max_path"], self.tn["min_sim"],
                                                   self.tn["min_length"], self.tn["max_iou"], _p(boxes), _p(nb), stream),
                       "tn_align")
            _lib.check(_lib.lib().vscb200_tn_box_scores(_p(sims), _p(d_so), _p(d_rl), n, _p(boxes), _p(nb), cap,
                                                        self.similarity_bias, _p(score), stream), "tn_box_scores")
        return boxes.cpu().numpy(), nb.cpu().numpy(), score.cpu().numpy()

    def _ts_arrays(self, videos, index):
        """(start, end) timestamp of every packed row, in pack order (VideoMetadata.get_timestamps, index.py:26-30)."""
        st = [np.asarray(v.timestamps) if np.asarray(v.timestamps).ndim == 1 else np.asarray(v.timestamps)[:, 0] for v in videos]
        en = [np.asarray(v.timestamps) if np.asarray(v.timestamps).ndim == 1 else np.asarray(v.timestamps)[:, 1] for v in videos]
        return np.concatenate(st), np.concatenate(en)

    def localize_all(self, candidates: Sequence) -> list:
        Match = _match_type()
        boxes, nb, box_scores = self.align(candidates)
        if not len(candidates) or not int(nb.sum()):
            return []
        if not hasattr(self, "_q_ts"):
            self._q_ts = self._ts_arrays(list(self.queries.values()), self.q_index)
            self._r_ts = self._ts_arrays(list(self.refs.values()), self.r_index)
        sel = np.arange(boxes.shape[1])[None, :] < nb[:, None]
        ci = np.nonzero(sel)[0]                                   # candidate of every box, candidate-major order
        bx = boxes[sel].astype(np.int64)
        qoff = np.array([self.q_index[c.query_id][0] for c in candidates], dtype=np.int64)[ci]
        roff = np.array([self.r_index[c.ref_id][0] for c in candidates], dtype=np.int64)[ci]
        qs, qe = self._q_ts[0][qoff + bx[:, 0]], self._q_ts[1][qoff + bx[:, 2]]
        rs, re = self._r_ts[0][roff + bx[:, 1]], self._r_ts[1][roff + bx[:, 3]]
        ms = box_scores[sel]
        fast = {VCSLLocalization.score: 0, VCSLLocalizationMaxSim.score: 1, VCSLLocalizationCandidateScore.score: 2}
        mode = fast.get(type(self).score, -1)
        matches = []
        for j, (i, a, b, c_, d_, m) in enumerate(zip(ci.tolist(), qs, qe, rs, re, ms)):
            c = candidates[i]
            if mode == 0:
                sc = 1.0
            elif mode == 1:
                sc = m
            elif mode == 2:
                sc = c.score
            else:
                match = Match(query_id=c.query_id, ref_id=c.ref_id, query_start=a, query_end=b, ref_start=c_, ref_end=d_,
                              score=0.0)
                sc = self.score(c, match, tuple(int(v) for v in bx[j]), m)
            matches.append(Match(query_id=c.query_id, ref_id=c.ref_id, query_start=a, query_end=b, ref_start=c_,
                                 ref_end=d_, score=sc))
        return matches

    def localize(self, candidate) -> list:
        return self.localize_all([candidate])

    def score(self, candidate, match, box, max_sim) -> float:
        """``max_sim``: sims[x1:x2, y1:y2].max() - similarity_bias of this box, computed on the device."""
        return 1.0


class VCSLLocalizationMaxSim(VCSLLocalization):        # localization.py:87-90
    def score(self, candidate, match, box, max_sim) -> float:
        return max_sim


class VCSLLocalizationCandidateScore(VCSLLocalization):   # localization.py:93-95
    def score(self, candidate, match, box, max_sim) -> float:
        return candidate.score
