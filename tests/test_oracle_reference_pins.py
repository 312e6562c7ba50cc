"""Container-only pins: the reference's OWN unit tests and functions, run unmodified from
/root/reference, over the oracle's faiss stand-in.  Skipped where /root/reference is absent
(the GPU box)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import refload

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not refload.available(), reason="/root/reference not present")


def _run_reference_unittests(faiss_dir):
    root = os.path.join(refload.D, "train/train_v106")
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([faiss_dir, os.path.join(REPO, "oracle", "refshim"), REPO, root])
    return subprocess.run([sys.executable, "-W", "ignore", "-m", "unittest", "discover", "-s", "tests", "-t", "."],
                          cwd=root, env=env, capture_output=True, text=True, timeout=600)


def test_reference_unit_tests_pass_on_oracle_faiss():
    """test_candidates.py:72-83 (exact 2.0/1.0/0.25), test_index.py:40-53, test_storage, test_metrics."""
    r = _run_reference_unittests(os.path.join(REPO, "oracle", "refshim"))
    tail = (r.stdout + r.stderr)[-1500:]
    assert r.returncode == 0, tail
    assert "OK" in tail, tail


def test_score_normalize_restated_equals_reference_function():
    refload.vsc_package("D_infer")
    try:
        from vsc.baseline.score_normalization import query_score_normalize, ref_score_normalize, score_normalize
        from vsc.index import VideoFeature
        from oracle import score_norm_np
        rng = np.random.default_rng(3)

        def vids(prefix, lens):
            return [VideoFeature(video_id=f"{prefix}{i}", timestamps=np.arange(n, dtype=np.float32),
                                 feature=rng.standard_normal((n, 32)).astype(np.float32)) for i, n in enumerate(lens)]

        q, r, z = vids("Q", [3, 5, 1]), vids("R", [4, 4, 6]), vids("N", [7, 9])
        cat = lambda vs: np.concatenate([v.feature for v in vs])
        for beta, nk in ((1.2, 1), (1.5, 4)):
            sq, sr = score_normalize(q, r, z, beta=beta, nk=nk)
            oq, orr, lvd = score_norm_np.score_normalize(cat(q), cat(r), cat(z), beta=beta, nk=nk)
            np.testing.assert_allclose(oq, cat(sq), atol=2e-7)
            np.testing.assert_allclose(orr, cat(sr), atol=2e-7)
        scores = {"Q0": 0.5, "Q1": 0.0001, "Q2": 0.9}
        sq = query_score_normalize(q, z, scores, low_var_dim=5, beta=1.2, nk=2)
        gated = np.concatenate([[scores[v.video_id] < 0.001] * len(v) for v in q])
        oq = score_norm_np.query_score_normalize(cat(q), cat(z), gated, low_var_dim_=5, beta=1.2, nk=2)
        np.testing.assert_allclose(oq, cat(sq), atol=2e-7)
        sr = ref_score_normalize(r, z)
        orr, _ = score_norm_np.ref_score_normalize(cat(r), cat(z))
        np.testing.assert_allclose(orr, cat(sr), atol=2e-7)
    finally:
        refload.unload_vsc()


def test_score_normalize_v2_restated_equals_reference_function():
    """The matching track's score_normalizev2 (M/vsc/baseline/score_normalization.py:115-156), run unmodified over the
    oracle's faiss stand-in, against oracle/score_norm_np.score_normalize_v2."""
    refload.vsc_package("M")
    try:
        from vsc.baseline.score_normalization import score_normalizev2
        from vsc.index import VideoFeature
        from oracle import score_norm_np
        rng = np.random.default_rng(4)

        def vids(prefix, lens):
            return [VideoFeature(video_id=f"{prefix}{i}", timestamps=np.arange(n, dtype=np.float32),
                                 feature=rng.standard_normal((n, 32)).astype(np.float32)) for i, n in enumerate(lens)]

        q, r, z = vids("Q", [3, 5, 1]), vids("R", [4, 4, 6]), vids("N", [17, 19])
        cat = lambda vs: np.concatenate([v.feature for v in vs])
        q0, r0, z0 = cat(q).copy(), cat(r).copy(), cat(z).copy()      # the reference adapts query.feature in place
        sq, sr = score_normalizev2(q, r, z, beta=0.35, nk=10)
        oq, orr = score_norm_np.score_normalize_v2(q0, r0, z0, beta=0.35, nk=10)
        np.testing.assert_allclose(oq, cat(sq), atol=3e-7)
        np.testing.assert_allclose(orr, cat(sr), atol=3e-7)
    finally:
        refload.unload_vsc()
