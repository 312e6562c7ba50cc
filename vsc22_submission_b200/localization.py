"""Frame-similarity matrices of all candidate video pairs on the device.

Mirrors ``LocalizationWithMetadata.similarity`` / ``VCSLLocalization.similarity`` and the list built in
``VCSLLocalization.localize_all`` (vsc/baseline/localization.py:27-66): for every ``CandidatePair`` the matrix
``queries[c.query_id].feature @ refs[c.ref_id].feature.T + similarity_bias``; in addition the per-row top-k that
``vcsl.vta.tn`` computes first (vta.py:262-265).  The reference does one numpy matmul per pair inside a 16-process
pool; here every video's descriptors are uploaded once and all pairs are computed by one launch.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Sequence, Tuple

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
