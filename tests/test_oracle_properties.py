"""Size-independent properties of the oracle restatements (hypothesis): they are what the GPU parity tests lean on at sizes
where no reference output exists."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import candidates_np, faiss_np, near_dup_np, resize_np, tn_np


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 40), st.integers(1, 50), st.integers(0, 2 ** 31 - 1), st.integers(2, 12), st.integers(1, 6))
def test_tn_boxes_are_valid_and_mutually_distinct(Q, R, seed, max_step, top_k):
    rng = np.random.default_rng(seed)
    s = rng.uniform(-0.3, 1.0, (Q, R)).astype(np.float32)
    boxes = tn_np.tn(s, tn_max_step=max_step, tn_top_k=top_k, min_length=1, max_iou=0.3)
    assert len(boxes) <= 11
    for i, (q0, r0, q1, r1) in enumerate(boxes):
        assert 0 <= q0 < q1 < Q and 0 <= r0 < r1 < R                      # extents exceed min_length = 1
        for other in boxes[:i]:
            assert tn_np._iou_max([q0, r0, q1, r1], [other]) < 0.3
    assert tn_np.tn(s * 0 - 1.0, tn_max_step=max_step, tn_top_k=top_k) == []   # nothing reaches min_sim


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 30), st.integers(1, 60), st.integers(0, 2 ** 31 - 1), st.integers(1, 200))
def test_global_topk_and_video_pairs(nq, nr, seed, k):
    rng = np.random.default_rng(seed)
    q, r = rng.standard_normal((nq, 8)).astype(np.float32), rng.standard_normal((nr, 8)).astype(np.float32)
    s, qi, ri = candidates_np.global_topk_pairs(q, r, k)
    assert len(s) == min(k, nq * nr) and (np.diff(s) <= 0).all()
    assert len(set(zip(qi.tolist(), ri.tolist()))) == len(s)
    S = candidates_np._scores(q, r, faiss_np.METRIC_INNER_PRODUCT)
    assert len(s) == nq * nr or S.reshape(-1)[np.setdiff1d(np.arange(nq * nr), qi * nr + ri)].max() <= s[-1]
    q_len = [nq // 2, nq - nq // 2] if nq > 1 else [1]
    r_len = [nr // 3, nr - nr // 3] if nr > 2 else [nr]
    cands = candidates_np.video_pair_candidates(s, qi, ri, q_len, r_len)
    assert len({(a, b) for a, b, _ in cands}) == len(cands) and all(x[2] >= y[2] for x, y in zip(cands, cands[1:]))
    assert cands[0][2] == s[0]
    # the threshold form returns exactly the pairs above the threshold
    t = float(np.median(S))
    ts, tq, tr = candidates_np.threshold_pairs(q, r, t)
    assert len(ts) == int((S > t).sum()) and (ts > t).all()


@settings(max_examples=25, deadline=None)
@given(st.integers(2, 40), st.integers(2, 40), st.integers(1, 48), st.integers(1, 48), st.integers(0, 2 ** 31 - 1))
def test_resize_range_identity_and_constants(h, w, oh, ow, seed):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    out = resize_np.resize_bicubic_u8(img, oh, ow)
    assert out.shape == (oh, ow, 3) and out.dtype == np.uint8
    np.testing.assert_array_equal(resize_np.resize_bicubic_u8(img, h, w), img)          # same size: untouched
    flat = np.full((h, w, 3), 137, np.uint8)
    assert np.abs(resize_np.resize_bicubic_u8(flat, oh, ow).astype(int) - 137).max() <= 1   # weights sum to 1 (+- quantisation)
    x = resize_np.preprocess(img, oh, ow, (0.5,) * 3, (0.5,) * 3)
    assert x.shape == (3, oh, ow) and x.dtype == np.float32 and x.min() >= -1.0 and x.max() <= 1.0


@settings(max_examples=30, deadline=None)
@given(st.integers(1, 60), st.integers(0, 2 ** 31 - 1))
def test_near_dup_filter_keeps_a_maximal_spread_subset(n, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, 16)).astype(np.float32)
    dup = rng.integers(0, n, n // 3)
    x[dup] = x[(dup + 1) % n] * rng.uniform(0.5, 2.0)                     # scaled copies: cosine 1
    keep = near_dup_np.keep_indices(x)
    assert len(keep) >= 1 and keep == sorted(keep)
    f = x / np.linalg.norm(x, axis=1, keepdims=True)
    sim = f[keep] @ f[keep].T - np.eye(len(keep))
    assert (sim <= 0.975 + 1e-6).all()                                    # no two survivors are near-duplicates
