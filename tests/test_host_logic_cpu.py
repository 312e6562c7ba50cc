"""Host-side logic of the reference-facing mirrors with the device calls stubbed out (runs without a GPU): grouping of
frame pairs into PairMatches, Match construction from boxes for both timestamp layouts, result ordering and limits."""
import dataclasses

import numpy as np
import torch

from vsc22_submission_b200 import candidates, localization, matching


@dataclasses.dataclass
class VF:
    video_id: str
    feature: np.ndarray
    timestamps: np.ndarray


@dataclasses.dataclass
class Cand:
    query_id: str
    ref_id: str
    score: float = 0.0


def _videos(prefix, lens, two_col):
    out = []
    for i, n in enumerate(lens):
        ts = np.arange(n, dtype=np.float32) * 0.5 + i
        out.append(VF(f"{prefix}{i}", np.zeros((n, 4), np.float32), np.stack([ts, ts + 0.5], 1) if two_col else ts))
    return out


def test_video_index_search_groups_pairs_in_first_appearance_order():
    queries, refs = _videos("Q", [3, 2], False), _videos("R", [2, 4, 1], True)
    vi = candidates.VideoIndex.__new__(candidates.VideoIndex)            # no device: fields set by hand
    vi.videos, vi._r_off = refs, candidates._offsets(refs)
    # (score, query row, bank row) best first: Q1/R1 twice, Q0/R0, Q1/R2, Q0/R1
    sc = torch.tensor([0.9, 0.8, 0.7, 0.6, 0.5])
    qi = torch.tensor([3, 4, 0, 4, 2])
    ri = torch.tensor([2, 5, 1, 6, 3])
    vi._frame_pairs = lambda q, k, thr=None: (sc, qi, ri)
    got = vi.search(queries, global_k=5)
    assert [(m.query_id, m.ref_id, len(m.matches)) for m in got] == [("Q1", "R1", 2), ("Q0", "R0", 1), ("Q1", "R2", 1),
                                                                     ("Q0", "R1", 1)]
    first = got[0].matches[0]                                            # query row 3 = Q1 frame 0; bank row 2 = R1 frame 0
    assert first.query_timestamps == (1.0, 1.0) and first.ref_timestamps == (1.0, 1.5) and abs(first.score - 0.9) < 1e-7
    assert got[0].matches[1].ref_timestamps == (2.5, 3.0)                # bank row 5 = R1 frame 3
    assert list(got[1].records())[0]["ref_end"] == 1.0                   # bank row 1 = R0 frame 1: (0.5, 1.0)

    class Sum:                                                           # any aggregation object works on the generic path
        def score(self, match):
            return candidates._CandidatePair(match.query_id, match.ref_id, sum(m.score for m in match.matches))

    cg = candidates.CandidateGeneration.__new__(candidates.CandidateGeneration)
    cg.aggregation, cg.index = Sum(), vi
    out = cg.query(queries, global_k=5)
    assert [(c.query_id, c.ref_id) for c in out] == [("Q1", "R1"), ("Q0", "R0"), ("Q1", "R2"), ("Q0", "R1")]
    assert abs(out[0].score - 1.7) < 1e-6 and len(cg.query(queries, global_k=5, limit=2)) == 2
    # MaxScoreAggregation takes the device reduction (stubbed here) and honours `limit`
    cg.aggregation = candidates.MaxScoreAggregation()
    vi.video_pairs = lambda q, k, thr=None: (np.array([0.9, 0.7, 0.6], np.float32), np.array([1, 0, 1]), np.array([1, 0, 2]))
    out = cg.query(queries, global_k=5, limit=2)
    assert [(c.query_id, c.ref_id, float(c.score)) for c in out] == [("Q1", "R1", np.float32(0.9)), ("Q0", "R0", np.float32(0.7))]
    assert candidates.MaxScoreAggregation().aggregate(got[0]) == sc[0].item()


def test_localize_all_builds_matches_from_boxes():
    for two_col in (False, True):
        queries, refs = _videos("Q", [6, 5], two_col), _videos("R", [7, 8], not two_col)
        cands = [Cand("Q1", "R0", 0.3), Cand("Q0", "R1", 0.6), Cand("Q0", "R0", 0.1)]
        boxes = np.zeros((3, 2, 4), np.int32)
        boxes[0, 0] = (0, 1, 4, 6)
        boxes[1, 0] = (1, 0, 5, 7)
        boxes[1, 1] = (2, 2, 3, 4)
        nb = np.array([1, 2, 0], np.int32)
        ms = np.array([[0.91, 0], [0.82, 0.73], [0, 0]], np.float32)
        for cls in (localization.VCSLLocalization, localization.VCSLLocalizationMaxSim,
                    localization.VCSLLocalizationCandidateScore):
            loc = cls.__new__(cls)
            loc.queries, loc.refs = {v.video_id: v for v in queries}, {v.video_id: v for v in refs}
            off = lambda vs: {v.video_id: (int(sum(len(u.feature) for u in vs[:i])), len(v.feature)) for i, v in enumerate(vs)}
            loc.q_index, loc.r_index = off(queries), off(refs)
            loc.align = lambda c: (boxes, nb, ms)
            got = loc.localize_all(cands)
            ts = lambda v, i, end: (v.timestamps[i] if v.timestamps.ndim == 1 else v.timestamps[i][1 if end else 0])
            want = []
            for i, c in enumerate(cands):
                for b in range(nb[i]):
                    x1, y1, x2, y2 = boxes[i, b]
                    score = {localization.VCSLLocalization: 1.0, localization.VCSLLocalizationMaxSim: ms[i, b],
                             localization.VCSLLocalizationCandidateScore: c.score}[cls]
                    want.append((c.query_id, c.ref_id, score, ts(loc.queries[c.query_id], x1, False),
                                 ts(loc.queries[c.query_id], x2, True), ts(loc.refs[c.ref_id], y1, False),
                                 ts(loc.refs[c.ref_id], y2, True)))
            assert [tuple(m) for m in got] == want
        # a subclass with its own score() goes through the per-match hook with the box and the device max
        class Custom(localization.VCSLLocalization):
            def score(self, candidate, match, box, max_sim):
                return box[2] - box[0] + float(max_sim)
        loc = Custom.__new__(Custom)
        loc.queries, loc.refs, loc.q_index, loc.r_index = ({v.video_id: v for v in queries}, {v.video_id: v for v in refs},
                                                           off(queries), off(refs))
        loc.align = lambda c: (boxes, nb, ms)
        assert [round(m.score, 2) for m in loc.localize_all(cands)] == [4.91, 4.82, 1.73]
        loc.align = lambda c: (boxes, np.zeros(3, np.int32), ms)
        assert loc.localize_all(cands) == []


def test_matching_batches_keep_the_reference_item_order():
    mf = matching.MatchingFeatures.__new__(matching.MatchingFeatures)
    cands = [("Q0", "R1", 0.5), ("Q2", "R0", 0.25)]
    imgs = torch.arange(2 * 2 * 3 * 3, dtype=torch.float32).reshape(2, 2, 3, 3)
    mf._images = lambda c, m, res, tr: (imgs if tr else imgs[:, :1], np.array([[1, 4, 3, 2], [0, 7, 3, 3]], np.int32))
    feats, infos = mf.classify_batch(cands, {}, (3, 3))
    assert tuple(feats.shape) == (4, 3, 3, 3) and infos == [["Q0", "R1", 0.5]] * 2 + [["Q2", "R0", 0.25]] * 2
    assert torch.equal(feats[1, 0], imgs[0, 1]) and torch.equal(feats[2, 2], imgs[1, 0])     # q@r.T, r@q.T, next pair ...
    f2, qids, rids, h, w, seg = mf.refine_batch(cands, {}, (3, 3))
    assert tuple(f2.shape) == (2, 3, 3, 3) and (qids, rids) == (["Q0", "Q2"], ["R1", "R0"])
    assert h.tolist() == [3, 3] and w.tolist() == [2, 3] and seg.tolist() == [1, 0]
