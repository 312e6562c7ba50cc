"""Oracle (B) against the committed golden vectors produced by the reference's own vsc package
(tests/golden/make_golden.py) -- CPU only."""
import os

import numpy as np

from oracle import faiss_np, score_norm_np


def _g(golden_dir):
    return np.load(os.path.join(golden_dir, "search_small.npz"))


def test_score_normalize_matches_reference(golden_dir):
    g = _g(golden_dir)
    q, r, lvd = score_norm_np.score_normalize(g["q_raw"], g["r_raw"], g["z_raw"], beta=1.2, nk=1)
    assert q.shape == g["sn_q"].shape and r.shape == g["sn_r"].shape
    np.testing.assert_allclose(q, g["sn_q"], rtol=0, atol=2e-7)
    np.testing.assert_allclose(r, g["sn_r"], rtol=0, atol=2e-7)
    # docstring identity (score_normalization.py:42-62): q'.r' = q.r + bias(q)
    s = q @ r.T
    s0 = q[:, :-1] @ r[:, :-1].T + q[:, -1:]
    np.testing.assert_allclose(s, s0, atol=1e-5)


def test_knn_and_range_match_reference(golden_dir):
    g = _g(golden_dir)
    index = faiss_np.index_factory(g["sn_r"].shape[1], "Flat", faiss_np.METRIC_INNER_PRODUCT)
    index.add(g["sn_r"])
    D, I = index.search(g["sn_q"], 10)
    np.testing.assert_array_equal(I, g["I10"])
    np.testing.assert_array_equal(D, g["D10"])
    lims, Dr, Ir = index.range_search(g["sn_q"], 0.0)
    np.testing.assert_array_equal(lims, g["lims"])
    np.testing.assert_array_equal(Ir, g["Ir"])
    np.testing.assert_array_equal(Dr, g["Dr"])
    assert lims.dtype == np.uint64 and I.dtype == np.int64 and D.dtype == np.float32


def test_search_semantics():
    rng = np.random.default_rng(0)
    xb = rng.standard_normal((50, 8)).astype(np.float32)
    xq = rng.standard_normal((7, 8)).astype(np.float32)
    for metric in (faiss_np.METRIC_INNER_PRODUCT, faiss_np.METRIC_L2):
        ix = faiss_np.IndexFlat(8, metric)
        ix.add(xb[:20]); ix.add(xb[20:])
        assert ix.ntotal == 50
        D, I = ix.search(xq, 60)          # k > ntotal -> padded
        assert (I[:, 50:] == -1).all()
        sign = -1 if metric == faiss_np.METRIC_INNER_PRODUCT else 1
        assert (np.diff(sign * D[:, :50].astype(np.float64), axis=1) >= 0).all()
        ref = xq @ xb.T if metric == faiss_np.METRIC_INNER_PRODUCT else ((xq[:, None] - xb[None]) ** 2).sum(-1)
        np.testing.assert_allclose(D[:, :50], np.take_along_axis(ref, I[:, :50], 1), rtol=1e-5, atol=1e-5)
        ix.reset()
        assert ix.ntotal == 0
        D, I = ix.search(xq, 3)
        assert (I == -1).all()
        lims, Dr, Ir = ix.range_search(xq, 0.0)
        assert lims.tolist() == [0] * 8 and len(Dr) == 0


def test_ties_go_to_lower_id():
    ix = faiss_np.IndexFlat(2, faiss_np.METRIC_INNER_PRODUCT)
    ix.add(np.array([[1, 0], [1, 0], [0, 1], [1, 0]], np.float32))
    D, I = ix.search(np.array([[1, 0]], np.float32), 3)
    assert I.tolist() == [[0, 1, 3]]
