"""Matching-track candidate features on the device (matching.MatchingFeatures -> csrc/pair_sims.cu) against the oracle
(oracle/matching_np.py, pinned to the reference functions in tests/test_oracle_matching.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_classify_and_refine_batches_match_oracle():
    from oracle import matching_np
    from matching_cases import make_case
    from vsc22_submission_b200.matching import MatchingFeatures
    for seed in (0, 1):
        query, ref, cands, len_map = make_case(seed, n_q=14, n_r=16, d=64)
        mf = MatchingFeatures(query, ref)
        feats, infos = mf.classify_batch(cands, len_map, (160, 160))
        want = matching_np.classify_images(query, ref, cands, len_map, (160, 160))
        assert tuple(feats.shape) == (2 * len(cands), 3, 160, 160) and feats.is_cuda
        got = feats.cpu().numpy()
        assert np.abs(got[:, 0] - want).max() <= 2e-6                 # exact-fp32 FFMA vs numpy sgemm
        assert (got[:, 0] == got[:, 1]).all() and (got[:, 0] == got[:, 2]).all()
        assert ((got[:, 0] == 0) == (want == 0)).all()                # identical padding
        assert infos == [[c[0], c[1], c[2]] for c in cands for _ in (0, 1)]
        f2, qids, rids, h, w, seg = mf.refine_batch(cands, len_map, (224, 224))
        w_img, w_seg, w_h, w_w = matching_np.refine_images(query, ref, cands, len_map, (224, 224))
        np.testing.assert_array_equal(seg, w_seg)                      # the same query copy is kept
        np.testing.assert_array_equal(h, w_h)
        np.testing.assert_array_equal(w, w_w)
        assert np.abs(f2.cpu().numpy()[:, 0] - w_img).max() <= 2e-6
        assert (qids, rids) == ([c[0] for c in cands], [c[1] for c in cands])
        assert len(set(w_seg.tolist())) > 1
    assert tuple(mf.classify_batch([], len_map)[0].shape) == (0, 3, 160, 160)
