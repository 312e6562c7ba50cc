"""oracle/candidates_np.py against (1) the golden candidate list the reference's own CandidateGeneration wrote
(tests/golden/search_small.npz), (2) the reference classes run live in the container, including a case large
enough to make range_search_max_results shrink its radius several times, and (3) the SURVEY 8d config-1
plumbing run: `python -m vsc.baseline.sscd_baseline` (eval.sh arguments) end to end -> candidates.csv / matches.csv."""
import os

import numpy as np
import pytest

from oracle import candidates_np, faiss_np, refload


def test_golden_candidates(golden_dir):
    g = np.load(os.path.join(golden_dir, "search_small.npz"))
    cands = candidates_np.candidates(g["sn_q"], g["sn_r"], g["q_len"], g["r_len"], int(g["global_k"]))
    assert [f"Q{a:06d}" for a, _, _ in cands] == list(g["cand_q"])
    assert [f"R{b:06d}" for _, b, _ in cands] == list(g["cand_r"])
    np.testing.assert_array_equal(np.array([s for _, _, s in cands], dtype=np.float32), g["cand_s"])


def test_global_topk_is_a_prefix_of_the_sorted_pairs():
    rng = np.random.default_rng(0)
    q, r = rng.standard_normal((17, 8)).astype(np.float32), rng.standard_normal((29, 8)).astype(np.float32)
    for metric in (faiss_np.METRIC_INNER_PRODUCT, faiss_np.METRIC_L2):
        s, qi, ri = candidates_np.global_topk_pairs(q, r, 50, metric)
        S = candidates_np._scores(q, r, metric)
        np.testing.assert_array_equal(s, S[qi, ri])
        ref = np.sort(S.reshape(-1))
        ref = ref[::-1][:50] if metric == faiss_np.METRIC_INNER_PRODUCT else ref[:50]
        np.testing.assert_array_equal(s, ref)
    s, qi, ri = candidates_np.threshold_pairs(q, r, 1.0)
    assert (s > 1.0).all() and len(s) == int((candidates_np._scores(q, r, 0) > 1.0).sum())
    assert (np.diff(s) <= 0).all()


def test_exact_ties_keep_row_order():
    q = np.eye(4, dtype=np.float32)[:2]
    r = np.concatenate([np.eye(4, dtype=np.float32)[:2]] * 3)          # every score is 0 or 1, many ties
    s, qi, ri = candidates_np.global_topk_pairs(q, r, 6)
    assert list(zip(qi, ri)) == [(0, 0), (0, 2), (0, 4), (1, 1), (1, 3), (1, 5)] and (s == 1).all()
    cands = candidates_np.video_pair_candidates(s, qi, ri, [1, 1], [2, 2, 2])
    assert [(a, b) for a, b, _ in cands] == [(0, 0), (0, 1), (0, 2), (1, 0), (1, 1), (1, 2)]


needs_ref = pytest.mark.skipif(not refload.available(), reason="/root/reference not present")


def _videos(VideoFeature, rng, prefix, n, lo, hi, d, base=0):
    out = []
    for i in range(n):
        m = int(rng.integers(lo, hi))
        f = rng.standard_normal((m, d)).astype(np.float32)
        f /= np.linalg.norm(f, axis=1, keepdims=True)
        ts = np.stack([np.arange(m), np.arange(m) + 1], 1).astype(np.float32)
        out.append(VideoFeature(video_id=f"{prefix}{base + i:06d}", timestamps=ts, feature=f))
    return out


@needs_ref
def test_reference_candidate_generation_equals_oracle():
    """300 query frames x 900 ref frames with global_k = 500: the reference emits 32 x 900 pairs in its first batch,
    then shrinks the radius on most of the following ones (exhaustive_search.py:262-267)."""
    refload.vsc_package("D_infer")
    try:
        from vsc.candidates import CandidateGeneration, MaxScoreAggregation
        from vsc.index import VideoFeature
        rng = np.random.default_rng(5)
        queries = _videos(VideoFeature, rng, "Q", 20, 8, 22, 32)
        refs = _videos(VideoFeature, rng, "R", 60, 8, 22, 32)
        for i in range(0, 20, 3):                        # planted copies
            n = min(len(queries[i]), len(refs[2 * i])) - 2
            queries[i].feature[:n] = refs[2 * i].feature[:n]
        cat = lambda vs: np.concatenate([v.feature for v in vs])
        lens = lambda vs: [len(v) for v in vs]
        for gk in (500, 37, 5000):
            ref_c = CandidateGeneration(refs, MaxScoreAggregation()).query(queries, global_k=gk)
            ora = candidates_np.candidates(cat(queries), cat(refs), lens(queries), lens(refs), gk)
            assert [(c.query_id, c.ref_id) for c in ref_c] == [(queries[a].video_id, refs[b].video_id) for a, b, _ in ora]
            np.testing.assert_array_equal(np.array([c.score for c in ref_c], np.float32),
                                          np.array([s for _, _, s in ora], np.float32))
    finally:
        refload.unload_vsc()


@needs_ref
def test_config1_plumbing_sscd_baseline_end_to_end(tmp_path):
    """SURVEY 8d config 1 (CPU part): store_features -> `sscd_baseline` with eval.sh's arguments (D/infer/eval.sh:12-16)
    -> candidates.csv / matches.csv exist; the candidate list equals the oracle's (first 25 per query,
    sscd_baseline.py:96-101) and the planted copies come out on top and are localised."""
    import pandas as pd
    refload.vsc_package("D_infer")
    try:
        from vsc.baseline import sscd_baseline
        from vsc.index import VideoFeature
        from vsc.storage import store_features
        rng = np.random.default_rng(7)
        queries = _videos(VideoFeature, rng, "Q", 4, 16, 17, 64, base=100000)
        refs = _videos(VideoFeature, rng, "R", 10, 20, 21, 64, base=200000)
        for i in range(3):
            f = refs[2 * i].feature[6:14] + 0.05 * rng.standard_normal((8, 64)).astype(np.float32)
            queries[i].feature[4:12] = f / np.linalg.norm(f, axis=1, keepdims=True)
        qf, rf, out = str(tmp_path / "q.npz"), str(tmp_path / "r.npz"), str(tmp_path / "outputs")
        store_features(qf, queries)
        store_features(rf, refs)
        args = sscd_baseline.parser.parse_args(["--query_features", qf, "--ref_features", rf, "--output_path", out,
                                                "--overwrite"])
        sscd_baseline.main(args)
        cand = pd.read_csv(os.path.join(out, "candidates.csv"))
        matches = pd.read_csv(os.path.join(out, "matches.csv"))
        assert list(cand.columns) == ["query_id", "ref_id", "score"]
        assert list(matches.columns) == ["query_id", "ref_id", "query_start", "query_end", "ref_start", "ref_end", "score"]
        cat = lambda vs: np.concatenate([v.feature for v in vs])
        lens = lambda vs: [len(v) for v in vs]
        ora = candidates_np.candidates(cat(queries), cat(refs), lens(queries), lens(refs), int(1200.0 * len(queries)))
        ora = ora[: int(25.0 * len(queries))]
        assert list(zip(cand.query_id, cand.ref_id)) == [(queries[a].video_id, refs[b].video_id) for a, b, _ in ora]
        np.testing.assert_allclose(cand.score.to_numpy(), [s for _, _, s in ora], rtol=1e-6)
        planted = {(queries[i].video_id, refs[2 * i].video_id) for i in range(3)}
        assert set(zip(cand.query_id[:3], cand.ref_id[:3])) == planted
        assert planted <= set(zip(matches.query_id, matches.ref_id))
    finally:
        refload.unload_vsc()
