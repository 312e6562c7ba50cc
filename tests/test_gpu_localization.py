"""Temporal-network alignment on the device (csrc/tn_align.cu through localization.VCSLLocalization*) against the
oracle restatement of vcsl.vta.tn (oracle/tn_np.py, pinned to the reference function in tests/test_oracle_tn.py)."""
import dataclasses

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


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


def _unit(x):
    return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)


def _dataset(seed, nq=24, nr=30, d=64):
    rng = np.random.default_rng(seed)
    refs = [VF(f"R{i}", _unit(rng.standard_normal((int(rng.integers(1, 90)), d))), None) for i in range(nr)]
    refs[1] = VF("R1", _unit(rng.standard_normal((3, d))), None)             # fewer reference frames than top-k
    queries = []
    for i in range(nq):
        n = int(rng.integers(1, 80))
        f = rng.standard_normal((n, d))
        for _ in range(int(rng.integers(0, 3))):                             # copied segments from reference i / i+1
            src = refs[(i + int(rng.integers(0, 2))) % nr].feature
            L = int(rng.integers(2, 30))
            L = min(L, n, len(src))
            a, b = int(rng.integers(0, n - L + 1)), int(rng.integers(0, len(src) - L + 1))
            f[a:a + L] = src[b:b + L] + 0.25 * rng.standard_normal((L, d)) / np.sqrt(d)
        queries.append(VF(f"Q{i}", _unit(f), None))
    for v in queries + refs:
        m = len(v.feature)
        two = len(v.video_id) % 2 == 1
        v.timestamps = np.stack([np.arange(m), np.arange(m) + 1], 1).astype(np.float32) if two else np.arange(m, dtype=np.float32)
    cands = [Cand(f"Q{i}", f"R{(i + j) % nr}", float(rng.uniform())) for i in range(nq) for j in range(3)]
    cands.append(Cand("Q0", "R1"))
    return queries, refs, cands


@pytest.mark.parametrize("kw,bias", [
    (dict(tn_max_step=5, min_length=4), 0.5),                     # sscd_baseline.py:113-121 (score-normalised run)
    (dict(tn_max_step=5, min_length=4), 0.0),                     # sscd_baseline.py:123-130
    (dict(), 0.0),                                                # TnVtaModel defaults (vta.py:503-504)
    (dict(tn_max_step=3, tn_top_k=2, min_length=1, max_path=3), 0.0),
    (dict(tn_max_step=16, tn_top_k=8, min_sim=0.5, max_iou=0.9, max_path=20), 0.0),
])
def test_tn_boxes_match_oracle(kw, bias):
    from oracle import tn_np
    from vsc22_submission_b200.localization import VCSLLocalizationMaxSim
    queries, refs, cands = _dataset(len(kw) + int(bias * 10))
    loc = VCSLLocalizationMaxSim(queries, refs, model_type="TN", similarity_bias=bias, concurrency=16, **kw)
    boxes, nb, score = loc.align(cands)
    sims = loc.similarities(cands)
    total = 0
    for i, (c, (key, s)) in enumerate(zip(cands, sims)):
        ref = tn_np.tn(s, **kw)
        got = boxes[i, :nb[i]].tolist()
        assert got == ref, (key, s.shape, got, ref)
        for b, (x1, y1, x2, y2) in enumerate(ref):
            assert score[i, b] == np.float32(s[x1:x2, y1:y2].max() - np.float32(bias))
        total += len(ref)
    assert total >= 10, total
    assert loc.align([])[1].shape == (0,)


def test_localize_all_builds_the_reference_matches():
    """localize_all == localization.py:52-76 evaluated on the oracle's boxes (Match fields, both timestamp layouts,
    MaxSim and CandidateScore scoring)."""
    from oracle import tn_np
    from vsc22_submission_b200.localization import VCSLLocalizationCandidateScore, VCSLLocalizationMaxSim
    queries, refs, cands = _dataset(77)
    kw = dict(tn_max_step=5, min_length=4)
    qd, rd = {v.video_id: v for v in queries}, {v.video_id: v for v in refs}
    ts = lambda v, i: (v.timestamps[i], v.timestamps[i]) if v.timestamps.ndim == 1 else tuple(v.timestamps[i])
    for cls, bias in ((VCSLLocalizationMaxSim, 0.5), (VCSLLocalizationCandidateScore, 0.0)):
        loc = cls(queries, refs, model_type="TN", similarity_bias=bias, concurrency=16, **kw)
        got = loc.localize_all(cands)
        want = []
        for c, (_, s) in zip(cands, loc.similarities(cands)):
            for x1, y1, x2, y2 in tn_np.tn(s, **kw):
                sc = s[x1:x2, y1:y2].max() - np.float32(bias) if cls is VCSLLocalizationMaxSim else c.score
                want.append((c.query_id, c.ref_id, sc, ts(qd[c.query_id], x1)[0], ts(qd[c.query_id], x2)[1],
                             ts(rd[c.ref_id], y1)[0], ts(rd[c.ref_id], y2)[1]))
        assert len(got) == len(want) > 5
        for g, w in zip(got, want):
            assert (g.query_id, g.ref_id, g.query_start, g.query_end, g.ref_start, g.ref_end) == (w[0], w[1], *w[3:])
            assert g.score == w[2]
        assert loc.localize(cands[0]) == [m for m in got if (m.query_id, m.ref_id) == ("Q0", "R0")]
    with pytest.raises(ValueError):
        VCSLLocalizationMaxSim(queries, refs, model_type="DTW")


def test_long_videos_and_many_pairs():
    """1 500 pairs of up to 600 x 700 frames (more pairs than resident warps, scratch sized by the longest query)."""
    from oracle import tn_np
    from vsc22_submission_b200.localization import VCSLLocalization
    rng = np.random.default_rng(5)
    d = 32
    refs = [VF(f"R{i}", _unit(rng.standard_normal((int(rng.integers(20, 700)), d))), None) for i in range(40)]
    queries = []
    for i in range(50):
        n = int(rng.integers(20, 600))
        f = rng.standard_normal((n, d))
        src = refs[i % 40].feature
        L = min(n, len(src), int(rng.integers(10, 200)))
        f[:L] = src[:L] + 0.3 * rng.standard_normal((L, d)) / np.sqrt(d)
        queries.append(VF(f"Q{i}", _unit(f), None))
    for v in queries + refs:
        v.timestamps = np.arange(len(v.feature), dtype=np.float32)
    cands = [Cand(f"Q{i}", f"R{(i + j) % 40}") for i in range(50) for j in range(30)]
    loc = VCSLLocalization(queries, refs, model_type="TN", tn_max_step=5, min_length=4)
    boxes, nb, _ = loc.align(cands)
    sims = loc.similarities(cands)
    for i in list(range(0, 1500, 30)) + list(range(7, 1500, 97)):       # the planted pairs and a sample of the rest
        assert boxes[i, :nb[i]].tolist() == tn_np.tn(sims[i][1], tn_max_step=5, min_length=4), i
    assert int((nb[::30] > 0).sum()) >= 40


def test_config1_pipeline_matches_reference_golden(golden_dir, tmp_path):
    """SURVEY 8d config 1 on the device: pipeline.run (= sscd_baseline.main, eval.sh arguments) on the fixture's features
    reproduces the candidates.csv / matches.csv the UNMODIFIED reference wrote for them (tests/golden/make_golden.py
    make_pipeline), without and with score normalisation."""
    import os

    import pandas as pd
    from vsc22_submission_b200 import pipeline
    g = np.load(os.path.join(golden_dir, "pipeline_small.npz"))

    def videos(prefix, base, arr, lens):
        out, i = [], 0
        for n, ln in enumerate(lens):
            ln = int(ln)
            ts = np.stack([np.arange(ln), np.arange(ln) + 1], 1).astype(np.float32)
            out.append(VF(f"{prefix}{base + n:06d}", arr[i:i + ln], ts))
            i += ln
        return out

    queries, refs = videos("Q", 100000, g["q"], g["q_len"]), videos("R", 200000, g["r"], g["r_len"])
    noise = videos("R", 300000, g["z"], g["z_len"])
    for tag, sn in (("plain", None), ("sn", noise)):
        out = str(tmp_path / tag)
        cf, mf = pipeline.run(queries, refs, out, score_norm_refs=sn, overwrite=True)
        cand, mt = pd.read_csv(cf), pd.read_csv(mf)
        assert list(cand.columns) == ["query_id", "ref_id", "score"]
        assert list(mt.columns) == ["query_id", "ref_id", "query_start", "query_end", "ref_start", "ref_end", "score"]
        ws = g[f"{tag}_cand_s"]
        np.testing.assert_allclose(cand.score.to_numpy(), ws, rtol=0, atol=2e-6)
        same = (cand.query_id.to_numpy(str) == g[f"{tag}_cand_q"]) & (cand.ref_id.to_numpy(str) == g[f"{tag}_cand_r"])
        near = np.zeros(len(ws), bool)
        close = np.abs(np.diff(ws)) <= 1e-6
        near[1:] |= close
        near[:-1] |= close
        assert (same | near).all() and same[:40].all()
        assert list(mt.query_id) == list(g[f"{tag}_match_q"]) and list(mt.ref_id) == list(g[f"{tag}_match_r"])
        np.testing.assert_array_equal(mt[["query_start", "query_end", "ref_start", "ref_end"]].to_numpy(), g[f"{tag}_match_box"])
        np.testing.assert_allclose(mt.score.to_numpy(), g[f"{tag}_match_s"], rtol=0, atol=5e-6)
    with pytest.raises(Exception, match="overwrite"):
        pipeline.run(queries, refs, str(tmp_path / "plain"))
