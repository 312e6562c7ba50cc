"""Parity of the CUDA similarity path (through the C ABI / faiss-compatible module) with the oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def faiss():
    from vsc22_submission_b200 import faiss_compat
    assert faiss_compat.get_num_gpus() > 0, "GPU tests need a CUDA device"
    return faiss_compat


def tie_aware_equal(I, D, Io, Do, gap=1e-6):
    """Indices must match wherever the oracle's neighbouring scores are separated by > gap
    (random data has no exact ties; near-ties may legitimately swap)."""
    ok = I == Io
    d = Do.astype(np.float64)
    near = np.zeros_like(ok)
    near[:, 1:] |= np.abs(np.diff(d, axis=1)) <= gap * np.maximum(1.0, np.abs(d[:, 1:]))
    near[:, :-1] |= np.abs(np.diff(d, axis=1)) <= gap * np.maximum(1.0, np.abs(d[:, :-1]))
    near[:, -1] = True      # the k-th entry may swap with the (k+1)-th
    return bool((ok | near).all()), float(ok.mean())


def test_golden_search_small(faiss, golden_dir):
    g = np.load(os.path.join(golden_dir, "search_small.npz"))
    ix = faiss.index_factory(g["sn_r"].shape[1], "Flat", faiss.METRIC_INNER_PRODUCT)
    ix.add(g["sn_r"])
    D, I = ix.search(g["sn_q"], 10)
    np.testing.assert_array_equal(I, g["I10"])
    np.testing.assert_allclose(D, g["D10"], rtol=1e-5, atol=1e-6)
    lims, Dr, Ir = ix.range_search(g["sn_q"], 0.0)
    np.testing.assert_array_equal(lims, g["lims"])
    np.testing.assert_array_equal(Ir, g["Ir"])
    np.testing.assert_allclose(Dr, g["Dr"], rtol=1e-5, atol=1e-6)
    assert lims.dtype == np.uint64 and I.dtype == np.int64 and D.dtype == np.float32


@pytest.mark.parametrize("nb,nq,d,k,metric", [(5, 3, 3, 2, 1), (1000, 37, 64, 10, 0), (20000, 256, 512, 10, 0),
                                               (4000, 64, 512, 1024, 0), (777, 33, 20, 5, 1), (3000, 100, 512, 1, 0),
                                               (9, 4, 512, 16, 0),
                                               # streaming form (nq <= 128 rows against a large bank, sim_stream.cu)
                                               (70000, 40, 512, 10, 0), (40000, 1, 64, 1, 1), (50000, 128, 512, 5, 0),
                                               (33000, 17, 40, 26, 1), (100001, 8, 512, 10, 0),
                                               # few rows x many columns, large k: two-level selection (select.cu)
                                               (70001, 5, 64, 1024, 0), (66000, 3, 32, 100, 1)])
def test_search_matches_oracle(faiss, nb, nq, d, k, metric):
    from oracle import faiss_np
    rng = np.random.default_rng(nb + nq)
    xb = rng.standard_normal((nb, d)).astype(np.float32)
    xq = rng.standard_normal((nq, d)).astype(np.float32)
    if metric == 0:
        xb /= np.linalg.norm(xb, axis=1, keepdims=True)
        xq /= np.linalg.norm(xq, axis=1, keepdims=True)
    a, b = faiss.IndexFlat(d, metric), faiss_np.IndexFlat(d, metric)
    a.add(xb[: nb // 3]); a.add(xb[nb // 3:]); b.add(xb)
    assert a.ntotal == nb
    D, I = a.search(xq, k)
    Do, Io = b.search(xq, k)
    valid = Io >= 0
    assert ((I >= 0) == valid).all()
    # scores within 1e-3 relative (north star); indices exact up to near-ties
    assert np.abs(D - Do)[valid].max() <= 1e-3 * max(1.0, np.abs(Do[valid]).max())
    ok, frac = tie_aware_equal(I[:, :min(k, nb)], D[:, :min(k, nb)], Io[:, :min(k, nb)], Do[:, :min(k, nb)])
    assert ok, f"index mismatch beyond near-ties (exact fraction {frac})"
    assert frac > 0.999
    if k > nb:
        assert (I[:, nb:] == -1).all()


def test_exact_ties_resolve_to_lower_id(faiss):
    ix = faiss.IndexFlat(4, faiss.METRIC_INNER_PRODUCT)
    xb = np.zeros((600, 4), np.float32)
    xb[:, 0] = 1.0
    xb[100, 0] = 2.0
    ix.add(xb)
    D, I = ix.search(np.array([[1, 0, 0, 0]], np.float32), 5)
    assert I.tolist() == [[100, 0, 1, 2, 3]] and D.tolist() == [[2.0, 1.0, 1.0, 1.0, 1.0]]


def test_streaming_exact_ties_resolve_to_lower_id(faiss):
    """Ties across and inside the 32-row groups of the streaming search."""
    ix = faiss.IndexFlat(8, faiss.METRIC_INNER_PRODUCT)
    xb = np.zeros((40000, 8), np.float32)
    xb[:, 0] = 1.0
    xb[33333, 0] = 2.0
    ix.add(xb)
    D, I = ix.search(np.array([[1, 0, 0, 0, 0, 0, 0, 0]], np.float32), 6)
    assert I.tolist() == [[33333, 0, 1, 2, 3, 4]] and D.tolist() == [[2.0, 1.0, 1.0, 1.0, 1.0, 1.0]]


@pytest.mark.parametrize("metric", [0, 1])
def test_range_search_matches_oracle(faiss, metric):
    from oracle import faiss_np
    rng = np.random.default_rng(5)
    xb = rng.standard_normal((6000, 64)).astype(np.float32)
    xq = rng.standard_normal((70, 64)).astype(np.float32)
    a, b = faiss.IndexFlat(64, metric), faiss_np.IndexFlat(64, metric)
    a.add(xb); b.add(xb)
    thr = 12.0 if metric == 0 else 90.0
    lims, D, I = a.range_search(xq, thr)
    lo, Do, Io = b.range_search(xq, thr)
    # strict threshold: hits whose oracle score is within 1e-5 of thr may differ
    S = b._scores(xq)
    border = np.abs(S - thr) <= 1e-4 * abs(thr)
    if not border.any():
        np.testing.assert_array_equal(lims, lo)
        np.testing.assert_array_equal(I, Io)
        np.testing.assert_allclose(D, Do, rtol=1e-5, atol=1e-5)
    for i in range(len(xq)):                      # ascending ids inside each row
        seg = I[int(lims[i]):int(lims[i + 1])]
        assert (np.diff(seg) > 0).all()
    # "everything passes" radius of _global_threshold_knn_search (index.py:146)
    lims, D, I = a.range_search(xq[:3], -1e10 if metric == 0 else 1e10)
    assert lims.tolist() == [0, 6000, 12000, 18000] and (I[:6000] == np.arange(6000)).all()


def test_fused_and_dense_topk_paths_agree(faiss, monkeypatch):
    """Small k on a big bank takes the fused-epilogue path; VSCB200_NO_FUSED_TOPK=1 forces the dense
    score block + radix select.  Both end in the same exact fp32 rescoring, so results are identical."""
    rng = np.random.default_rng(21)
    xb = rng.standard_normal((30000, 128)).astype(np.float32)
    xb[5000:5040] = xb[77]                       # 41 exact duplicates: ties must resolve to the lower ids
    xq = rng.standard_normal((300, 128)).astype(np.float32)
    xq[0] = xb[77]
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("VSCB200_NO_FUSED_TOPK", mode)
        for metric in (faiss.METRIC_INNER_PRODUCT, faiss.METRIC_L2):
            ix = faiss.IndexFlat(128, metric)
            ix.add(xb)
            res[mode, metric] = ix.search(xq, 10)
    for metric in (faiss.METRIC_INNER_PRODUCT, faiss.METRIC_L2):
        (D0, I0), (D1, I1) = res["0", metric], res["1", metric]
        np.testing.assert_array_equal(I0, I1)
        np.testing.assert_array_equal(D0, D1)
    I = res["0", faiss.METRIC_INNER_PRODUCT][1]
    assert I[0].tolist() == [77] + list(range(5000, 5009))


def test_empty_and_reset(faiss):
    ix = faiss.IndexFlat(8, faiss.METRIC_INNER_PRODUCT)
    q = np.ones((2, 8), np.float32)
    D, I = ix.search(q, 3)
    assert (I == -1).all() and (D < -1e38).all()
    lims, Dr, Ir = ix.range_search(q, 0.0)
    assert lims.tolist() == [0, 0, 0] and len(Dr) == 0
    ix.add(np.eye(8, dtype=np.float32))
    assert ix.ntotal == 8
    D, I = ix.search(q[:0], 3)
    assert D.shape == (0, 3)
    ix.reset()
    assert ix.ntotal == 0
    ix.add(2 * np.eye(8, dtype=np.float32)[:2])
    D, I = ix.search(q, 1)
    assert I.ravel().tolist() == [0, 0] and D.ravel().tolist() == [2.0, 2.0]


def test_full_size_properties(faiss):
    """BASELINE config 3 sizes (10k x 40k x 512): size-independent properties instead of the oracle --
    self-search returns the row itself first with score ~1, scores are sorted, the planted copies are
    found, and a sampled slice agrees with the oracle."""
    from oracle import faiss_np
    rng = np.random.default_rng(3)
    R = rng.standard_normal((40000, 512)).astype(np.float32)
    R /= np.linalg.norm(R, axis=1, keepdims=True)
    Q = rng.standard_normal((10000, 512)).astype(np.float32)
    Q /= np.linalg.norm(Q, axis=1, keepdims=True)
    planted = rng.integers(0, 40000, size=500)
    Q[:500] = R[planted]
    ix = faiss.IndexFlat(512, faiss.METRIC_INNER_PRODUCT)
    ix.add(R)
    D, I = ix.search(Q, 10)
    assert (np.diff(D, axis=1) <= 0).all()
    assert (I[:500, 0] == planted).all() and np.allclose(D[:500, 0], 1.0, atol=1e-5)
    ob = faiss_np.IndexFlat(512, faiss_np.METRIC_INNER_PRODUCT)
    ob.add(R)
    sel = rng.choice(10000, 64, replace=False)
    Do, Io = ob.search(Q[sel], 10)
    ok, frac = tie_aware_equal(I[sel], D[sel], Io, Do)
    assert ok and frac > 0.995
    assert np.abs(D[sel] - Do).max() < 1e-5


def test_simt_and_tensor_core_paths_agree(faiss, monkeypatch):
    """The exact-fp32 SIMT kernel (VSCB200_FORCE_SIMT=1) and the split-bf16 tcgen05 kernel give the
    same neighbours and scores to fp32 rounding."""
    rng = np.random.default_rng(11)
    xb = rng.standard_normal((5000, 96)).astype(np.float32)
    xq = rng.standard_normal((130, 96)).astype(np.float32)
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("VSCB200_FORCE_SIMT", mode)
        ix = faiss.IndexFlat(96, faiss.METRIC_INNER_PRODUCT)
        ix.add(xb)
        res[mode] = ix.search(xq, 20)
    (D0, I0), (D1, I1) = res["0"], res["1"]
    assert np.abs(D0 - D1).max() < 2e-5
    ok, frac = tie_aware_equal(I0, D0, I1, D1, gap=2e-6)
    assert ok and frac > 0.999


def test_dense_scores_fp32_equivalent(faiss):
    import torch
    from vsc22_submission_b200 import search
    g = torch.Generator(device="cuda").manual_seed(0)
    unit = lambda n, d: torch.nn.functional.normalize(torch.randn((n, d), generator=g, device="cuda"))
    for (nq, nr, d) in [(700, 3001, 512), (65, 130, 24)]:
        Q, R = unit(nq, d), unit(nr, d)
        ix = search.DeviceIndex(d)
        ix.add(R)
        S = ix.scores(Q)
        ref = Q.double() @ R.double().T
        assert S.shape == (nq, nr)
        # dense tensor-core scores: 2-way bf16 split, ~2e-7 rms / <2e-6 max on unit vectors (search()
        # rescoring makes the returned top-k scores exact fp32; this is the raw score block)
        # error scales with |q||r| (= 1 here): 2^-18-level operand representation, <= ~8e-6 worst case
        assert (S.double() - ref).abs().max().item() < 1e-5
        assert (S.double() - ref).pow(2).mean().sqrt().item() < 2e-6


def test_ensemble_pca_tail_matches_oracle():
    """per-model normalize -> concat -> PCA.transform (concat_pca_sn.py:56-61) on the device."""
    import torch
    from oracle import pca_np
    from vsc22_submission_b200.ensemble import B200PCA
    rng = np.random.default_rng(4)
    dims = (512, 512, 512, 512)
    parts = [rng.standard_normal((1000, d)).astype(np.float32) * (i + 1) for i, d in enumerate(dims)]
    parts[2][5] = 0.0
    mean = (rng.standard_normal(sum(dims)) * 0.01).astype(np.float32)
    comp = np.linalg.qr(rng.standard_normal((sum(dims), 512)))[0].T.astype(np.float32)
    ref = pca_np.ensemble_pca(parts, mean, comp)
    pca = B200PCA(mean, comp)
    got = pca.transform_parts([torch.from_numpy(p).cuda() for p in parts]).cpu().numpy()
    assert np.abs(got - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())       # 1000 frames: split-bf16 tcgen05 projection
    assert np.abs(pca.transform_parts_host(parts) - got).max() == 0.0
    few = pca.transform_parts([torch.from_numpy(p[:100]).cuda() for p in parts]).cpu().numpy()     # < 256 frames: fp32 FFMA kernel
    assert np.abs(few - ref[:100]).max() <= 1e-5 * max(1.0, np.abs(ref).max())
    assert pca.transform_parts([torch.zeros((0, d), device="cuda") for d in dims]).shape == (0, 512)


def test_pair_similarity_matrices_and_topk():
    """per-candidate-pair `a @ b.T + bias` (localization.py:32-35,52-57) and the per-row top-k that
    opens vcsl.vta.tn (vta.py:262-265), all pairs in one launch, against numpy."""
    import dataclasses
    from vsc22_submission_b200.localization import PairSimilarity

    @dataclasses.dataclass
    class VF:
        video_id: str
        feature: np.ndarray

    @dataclasses.dataclass
    class Cand:
        query_id: str
        ref_id: str

    rng = np.random.default_rng(9)
    qs = [VF(f"Q{i}", rng.standard_normal((int(rng.integers(1, 150)), 512)).astype(np.float32)) for i in range(12)]
    rs = [VF(f"R{i}", rng.standard_normal((int(rng.integers(1, 200)), 512)).astype(np.float32)) for i in range(15)]
    rs[3] = VF("R3", rng.standard_normal((3, 512)).astype(np.float32))          # fewer reference frames than k
    cands = [Cand(f"Q{int(rng.integers(12))}", f"R{int(rng.integers(15))}") for _ in range(60)] + [Cand("Q0", "R3")]
    ps = PairSimilarity(qs, rs, similarity_bias=0.5)
    qd, rd = {v.video_id: v.feature for v in qs}, {v.video_id: v.feature for v in rs}
    got = ps.similarities_topk(cands, top_k=5)
    assert [g[0] for g in got] == [f"{c.query_id}-{c.ref_id}" for c in cands]
    for c, (_, sims, ti, tv) in zip(cands, got):
        ref = np.matmul(qd[c.query_id], rd[c.ref_id].T) + 0.5
        assert sims.shape == ref.shape and np.abs(sims - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max())
        top = min(5, ref.shape[1])
        order = np.argsort(-sims, kind="stable")[:, :top]                     # on OUR fp32 values: exact expectation
        np.testing.assert_array_equal(ti, order)
        np.testing.assert_array_equal(tv, np.take_along_axis(sims, order, axis=-1))
    plain = ps.similarities(cands[:3])
    assert all(np.array_equal(a[1], b[1]) for a, b in zip(plain, got[:3])) and ps.similarities([]) == []


def _videos(prefix, arr, lens):
    import dataclasses

    @dataclasses.dataclass
    class VF:
        video_id: str
        feature: np.ndarray
        timestamps: np.ndarray = None

    out, i = [], 0
    for n, ln in enumerate(lens):
        out.append(VF(f"{prefix}{n:06d}", arr[i:i + int(ln)]))
        i += int(ln)
    return out


def test_score_normalize_matches_reference_golden(golden_dir):
    """search.score_normalize (device) == the reference's own score_normalize on the golden descriptors
    (score_normalization.py:33-104, fixture made by tests/golden/make_golden.py through the unmodified vsc package)."""
    from vsc22_submission_b200 import search
    g = np.load(os.path.join(golden_dir, "search_small.npz"))
    q, r, z = (_videos(p, g[k], g[l]) for p, k, l in (("Q", "q_raw", "q_len"), ("R", "r_raw", "r_len"), ("N", "z_raw", "z_len")))
    sq, sr = search.score_normalize(q, r, z, beta=1.2)
    got_q, got_r = np.concatenate([v.feature for v in sq]), np.concatenate([v.feature for v in sr])
    np.testing.assert_allclose(got_q, g["sn_q"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(got_r, g["sn_r"], rtol=2e-5, atol=2e-6)
    assert [v.video_id for v in sq] == [v.video_id for v in q] and sq[0].feature.shape == q[0].feature.shape
    with pytest.raises(Exception, match="against VSC rules"):
        search.score_normalize(q, r, r, beta=1.2)


def test_query_and_ref_score_normalize_match_oracle():
    """query_score_normalize (video-score gate, nk > 1, fixed low_var_dim) and ref_score_normalize
    (score_normalization.py:107-192) against the numpy restatement; a bank large enough for the tensor-core path."""
    from oracle import score_norm_np
    from vsc22_submission_b200 import search
    rng = np.random.default_rng(12)
    d = 64
    qa, za = rng.standard_normal((90, d)).astype(np.float32), rng.standard_normal((5000, d)).astype(np.float32)
    ra = rng.standard_normal((300, d)).astype(np.float32)
    q, z, r = _videos("Q", qa, [30, 40, 20]), _videos("N", za, [2500, 2500]), _videos("R", ra, [100, 200])
    scores = {"Q000000": 0.5, "Q000001": 0.0, "Q000002": 0.9}
    gated = np.concatenate([np.full(30, False), np.full(40, True), np.full(20, False)])
    got = np.concatenate([v.feature for v in search.query_score_normalize(q, z, scores, low_var_dim=5, beta=1.2, nk=2)])
    ref = score_norm_np.query_score_normalize(qa, za, gated, low_var_dim_=5, beta=1.2, nk=2)
    np.testing.assert_allclose(got, ref, rtol=2e-5, atol=2e-5)
    assert (got[30:70, -1] == -100.0).all()
    got_r = np.concatenate([v.feature for v in search.ref_score_normalize(r, z)])
    ref_r, _ = score_norm_np.ref_score_normalize(ra, za)
    np.testing.assert_allclose(got_r, ref_r, rtol=2e-5, atol=2e-6)
    # matching-track signature (M/infer/infer_matching.py:213): (queries, refs, low_var_dim, beta=1.5, nk=10), no gate
    got_m = np.concatenate([v.feature for v in search.query_score_normalize(q, z, 7, beta=1.5, nk=10)])
    ref_m = score_norm_np.query_score_normalize(qa, za, np.zeros(90, bool), low_var_dim_=7, beta=1.5, nk=10)
    np.testing.assert_allclose(got_m, ref_m, rtol=2e-5, atol=2e-5)


def test_near_duplicate_frame_filter_and_query_tail():
    """ensemble.near_dup_keep / query_tail (extract_query_feats.py:169-204) against the oracle (pinned to the reference's
    own source lines in tests/test_oracle_near_dup.py)."""
    import torch
    from oracle import near_dup_np, pca_np
    from test_oracle_near_dup import videos
    from vsc22_submission_b200.ensemble import B200PCA, near_dup_keep, query_tail
    vids = videos()
    for x in vids[:-1]:
        keep = near_dup_keep(torch.from_numpy(x).cuda()).cpu().numpy()
        assert np.flatnonzero(keep).tolist() == near_dup_np.keep_indices(x)
    keep = np.flatnonzero(near_dup_keep(torch.from_numpy(vids[-1]).cuda()).cpu().numpy())
    assert len(keep) == 3 and len({k // 5 for k in keep}) == 3            # exact duplicates: one frame per group survives
    assert near_dup_keep(torch.zeros((0, 8), device="cuda")).shape == (0,)
    z = vids[2].copy()
    z[3] = 0.0                                           # an all-zero descriptor: NaN similarities, must not crash
    assert near_dup_keep(torch.from_numpy(z).cuda()).shape == (len(z),)
    rng = np.random.default_rng(8)
    parts = [rng.standard_normal((40, 64)).astype(np.float32) * (i + 1) for i in range(4)]
    for p in parts:
        p[10:14] = p[9] + 0.01 * rng.standard_normal((4, 64)).astype(np.float32)   # a static shot
    mean = (rng.standard_normal(256) * 0.01).astype(np.float32)
    comp = np.linalg.qr(rng.standard_normal((256, 32)))[0].T.astype(np.float32)
    feats, idx = query_tail(B200PCA(mean, comp), [torch.from_numpy(p).cuda() for p in parts])
    cat = np.concatenate([p / np.linalg.norm(p, axis=1, keepdims=True) for p in parts], axis=1)
    want_idx = near_dup_np.keep_indices(cat)
    assert idx.cpu().tolist() == want_idx and len(want_idx) < 40
    ref = pca_np.ensemble_pca([p[want_idx] for p in parts], mean, comp)
    assert np.abs(feats.cpu().numpy() - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())


def test_single_pass_search_equals_split_path_and_oracle(faiss, monkeypatch):
    """k <= 10 on a large bank: selection on ONE bf16 MMA per product + error margin + exact rescoring (sim_tc1.cu) must
    return exactly what the split-bf16 (3 MMAs) path and the fp32 oracle return -- also where the margin is crowded:
    300 bank rows within 2e-3 of a query (the row buffers spill into the query's global list) and 700 such rows (the
    list overflows -> exhaustive fp32 fallback for that query)."""
    from oracle import faiss_np
    rng = np.random.default_rng(31)
    d = 256
    unit = lambda x: (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)
    xb = unit(rng.standard_normal((50000, d)))
    xq = unit(rng.standard_normal((700, d)))
    xb[20000:20300] = unit(xq[5] + 2e-3 * rng.standard_normal((300, d)))        # crowded margin for query 5
    xb[41000:41012] = unit(xq[9] + 1e-2 * rng.standard_normal((12, d)))         # a dozen close rows for query 9
    xb[30000:30700] = unit(xq[7] + 2e-3 * rng.standard_normal((700, d)))        # more than a candidate list holds: query 7 is redone exhaustively
    xb[45000:45700] = unit(xq[650] + 1e-3 * rng.standard_normal((700, d)))      # ... and one in the last query tile
    res = {}
    for passes in ("1", "3"):
        monkeypatch.setenv("VSCB200_SIM_PASSES", passes)
        for metric in (faiss.METRIC_INNER_PRODUCT, faiss.METRIC_L2):
            ix = faiss.IndexFlat(d, metric)
            ix.add(xb)
            res[passes, metric] = ix.search(xq, 10)
            if passes == "1":       # queries 7 and 650 must have gone through the exhaustive kernel
                from vsc22_submission_b200 import _lib
                assert _lib.lib().vscb200_index_last_fallbacks(ix._h.ptr) >= 2
    for metric in (faiss.METRIC_INNER_PRODUCT, faiss.METRIC_L2):
        (D1, I1), (D3, I3) = res["1", metric], res["3", metric]
        np.testing.assert_array_equal(I1, I3)
        np.testing.assert_array_equal(D1, D3)          # both report the one exact fp32 summation order
        o = faiss_np.IndexFlat(d, metric)
        o.add(xb)
        Do, Io = o.search(xq, 10)
        ok, frac = tie_aware_equal(I1, D1, Io, Do)
        assert ok and frac > 0.995, frac
        assert np.abs(D1 - Do).max() <= 1e-5


def test_single_pass_search_ragged_shapes(faiss):
    """Bank / query counts that are not multiples of the 256-wide pair tiles, several query blocks, k = 1."""
    from oracle import faiss_np
    rng = np.random.default_rng(32)
    for (nb, nq, d, k) in [(2049, 129, 64, 1), (70001, 1000, 512, 10), (5000, 513, 96, 7)]:
        xb = rng.standard_normal((nb, d)).astype(np.float32)
        xq = rng.standard_normal((nq, d)).astype(np.float32)
        a, b = faiss.IndexFlat(d, faiss.METRIC_INNER_PRODUCT), faiss_np.IndexFlat(d, faiss.METRIC_INNER_PRODUCT)
        a.add(xb); b.add(xb)
        D, I = a.search(xq, k)
        Do, Io = b.search(xq, k)
        ok, frac = tie_aware_equal(I, D, Io, Do)
        assert ok and frac > 0.999, (nb, nq, d, k, frac)
        assert np.abs(D - Do).max() <= 1e-3 * np.abs(Do).max()


def test_score_normalize_v2_matches_oracle(faiss):
    """Matching-track score_normalizev2 (M/vsc/baseline/score_normalization.py:115-156) on the device: noise search through
    the single-pass top-k path (nk = 10), gather-mean-subtract-normalise kernel; against the numpy oracle (pinned to the
    reference function in tests/test_oracle_reference_pins.py), array form and the list-of-videos mirror."""
    import dataclasses
    import torch
    from oracle import score_norm_np
    from vsc22_submission_b200 import search
    rng = np.random.default_rng(41)
    q = rng.standard_normal((300, 64)).astype(np.float32)
    r = rng.standard_normal((500, 64)).astype(np.float32)
    z = rng.standard_normal((3000, 64)).astype(np.float32)
    oq, orr = score_norm_np.score_normalize_v2(q, r, z, beta=0.35, nk=10)
    dq, dr = search.score_normalize_v2_tensors(torch.from_numpy(q).cuda(), torch.from_numpy(r).cuda(), torch.from_numpy(z).cuda(),
                                               beta=0.35, nk=10)
    assert np.abs(dq.cpu().numpy() - oq).max() <= 2e-6 and np.abs(dr.cpu().numpy() - orr).max() <= 2e-6

    @dataclasses.dataclass
    class VF:
        video_id: str
        feature: np.ndarray

    vids = lambda pre, x, per: [VF(f"{pre}{i}", x[i:i + per].copy()) for i in range(0, x.shape[0], per)]
    sq, sr = search.score_normalizev2(vids("Q", q, 50), vids("R", r, 100), vids("N", z, 500), beta=0.35, nk=10)
    assert np.abs(np.concatenate([v.feature for v in sq]) - oq).max() <= 2e-6
    assert [v.video_id for v in sr] == [f"R{i}" for i in range(0, 500, 100)]
    with pytest.raises(Exception, match="against VSC rules"):
        search.score_normalizev2(vids("Q", q, 50), vids("R", r, 100), vids("R", z, 500))


def test_sharded_step_primitives_single_process(faiss):
    """The two collectives of the sharded config-3 step, emulated in one process: (1) per-shard column moments summed over
    the shards -> the same low-variance column as numpy on the whole bank, also for near-tied columns; (2) two partial
    top-k results side by side in one key tensor, merged per column range, with the per-row bias added after the merge."""
    import torch
    from vsc22_submission_b200 import search
    rng = np.random.default_rng(51)
    z = rng.standard_normal((5000, 96)).astype(np.float32)
    z[:, 17] *= 0.5
    z[:, 63] = z[:, 17] * (1 + 1e-4) + 3.0                      # nearly tied variances, very different means
    shards = [z[:1800], z[1800:1801], z[1801:]]
    m3 = sum(search.col_moments_local(torch.from_numpy(s_).cuda()) for s_ in shards)
    got = int(search.var_argmin_moments(m3, z.shape[0]).item())
    assert got == int(z.astype(np.float64).var(axis=0).argmin()) == 17
    # merge with column ranges and bias
    nq, k = 700, 10
    parts = 3
    Dn = rng.standard_normal((parts, nq, 1)).astype(np.float32)
    Dr = rng.standard_normal((parts, nq, k)).astype(np.float32)
    Dr.sort(axis=2); Dr = Dr[:, :, ::-1].copy()
    Ir = np.stack([rng.permutation(1000)[:k][None].repeat(nq, 0) + 1000 * p_ for p_ in range(parts)]).astype(np.int64)
    keys = torch.stack([torch.cat([search.pack_topk(torch.from_numpy(Dn[p_]).cuda(), torch.full((nq, 1), 7 + p_, dtype=torch.int64, device="cuda")),
                                   search.pack_topk(torch.from_numpy(Dr[p_]).cuda(), torch.from_numpy(Ir[p_]).cuda())], dim=1)
                        for p_ in range(parts)]).contiguous()
    Dz, _ = search.merge_packed_topk_cols(keys, 0, 1, 1)
    np.testing.assert_array_equal(Dz.cpu().numpy()[:, 0], Dn.max(axis=0)[:, 0])
    bias = torch.from_numpy(rng.standard_normal(nq).astype(np.float32)).cuda()
    D, I = search.merge_packed_topk_cols(keys, 1, k, k, bias=bias)
    allD = np.concatenate(list(Dr), axis=1); allI = np.concatenate(list(Ir), axis=1)
    order = np.argsort(-allD, axis=1, kind="stable")[:, :k]
    np.testing.assert_array_equal(I.cpu().numpy(), np.take_along_axis(allI, order, 1))
    np.testing.assert_array_equal(D.cpu().numpy(), np.take_along_axis(allD, order, 1) + bias.cpu().numpy()[:, None])


@pytest.mark.gpu
@pytest.mark.parametrize("d,drop,fill,use_bias", [(512, "dev", 0.0, False), (512, 7, 1.0, False), (64, -1, 0.0, True),
                                                  (20, 0, 1.0, True)])
def test_add_sn_stores_what_sn_transform_plus_add_stores(faiss, d, drop, fill, use_bias):
    """DeviceIndex.add_sn (one pass over the raw rows) vs sn_transform + add: identical stored rows, identical search
    results (scores bit for bit: same planes, same norms, same margins)."""
    import torch
    from vsc22_submission_b200 import search
    g = torch.Generator(device="cuda").manual_seed(d)
    x = torch.randn((5000, d), generator=g, device="cuda")
    x[17] = 0.0                                             # sklearn.normalize leaves zero rows zero
    q = torch.nn.functional.normalize(torch.randn((300, d), generator=g, device="cuda"))
    bias = torch.randn((5000,), generator=g, device="cuda") if use_bias else None
    dd = search.low_var_dim_device(x) if drop == "dev" else drop
    t = search.sn_transform(x, dd, True, fill=fill, bias=bias)
    a = search.DeviceIndex(t.shape[1]); a.add(t[:2000]); a.add(t[2000:])
    b = search.DeviceIndex(t.shape[1])
    b.add_sn(x[:2000], dd, True, fill=fill, bias=None if bias is None else bias[:2000])
    b.add_sn(x[2000:], dd, True, fill=fill, bias=None if bias is None else bias[2000:])
    assert torch.equal(a.reconstruct_n(0, 5000), b.reconstruct_n(0, 5000))
    qq = torch.nn.functional.normalize(torch.randn((300, t.shape[1]), generator=g, device="cuda"))
    for k in (1, 10, 64):
        Da, Ia = a.search(qq, k)
        Db, Ib = b.search(qq, k)
        assert torch.equal(Ia, Ib) and torch.equal(Da, Db)


@pytest.mark.gpu
def test_score_normalized_search_equals_the_two_step_form(faiss):
    import torch
    from vsc22_submission_b200 import search
    g = torch.Generator(device="cuda").manual_seed(3)
    unit = lambda n: torch.nn.functional.normalize(torch.randn((n, 512), generator=g, device="cuda"))
    q, r, z = unit(700), unit(9000), unit(6000)
    D, I = search.score_normalized_search(q, r, z, 10, beta=1.2, nk=1)
    q_t, r_t, _ = search.score_normalize_tensors(q, r, z, beta=1.2, nk=1)
    ix = search.DeviceIndex(r_t.shape[1]); ix.add(r_t)
    D2, I2 = ix.search(q_t, 10)
    assert torch.equal(I, I2) and torch.equal(D, D2)


@pytest.mark.gpu
def test_pack_topk_into_columns_equals_pack_topk(faiss):
    """The exchange layout of the sharded similarity step: several partial results side by side, each packed straight into its
    columns (vscb200_topk_pack_cols) -- the same keys as pack_topk + a strided copy, -1 ids as padding."""
    import torch
    from vsc22_submission_b200 import search, sharding
    g = torch.Generator(device="cuda").manual_seed(9)
    nq = 1000
    parts = []
    for k in (1, 10, 3):
        D = torch.randn((nq, k), generator=g, device="cuda")
        I = torch.randint(0, 1 << 31, (nq, k), generator=g, device="cuda", dtype=torch.int64)
        I[::7, -1] = -1
        parts.append((D, I))
    keys = sharding.gather_partial_topk_multi(parts)          # world size 1: [1, nq, 14]
    assert keys.shape == (1, nq, 14) and keys.dtype == torch.int64
    want = torch.cat([search.pack_topk(D, I) for D, I in parts], dim=1)
    assert torch.equal(keys[0], want)
    D10, I10 = search.merge_packed_topk_cols(keys, 1, 10, 10)
    order = torch.argsort(parts[1][0].masked_fill(parts[1][1] < 0, float("-inf")), dim=1, descending=True, stable=True)
    assert torch.equal(I10, torch.gather(parts[1][1], 1, order))
