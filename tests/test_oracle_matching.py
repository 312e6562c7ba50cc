"""oracle/matching_np.py == the reference's generate_candidates_classfiy_feature / generate_matching_feature +
MatchClassifyDataset / MatchRefineDataset (imported from /root/reference; container-only)."""
import importlib.util
import os
import sys

import numpy as np
import pytest

from matching_cases import make_case
from oracle import matching_np, refload

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_select_segment_basics():
    q = np.eye(6, dtype=np.float32)
    r = np.eye(6, dtype=np.float32)[3:]
    k, kept = matching_np.select_segment(q, r, 3)
    assert k == 1 and kept.shape == (3, 6)
    assert matching_np.select_segment(q, r, 6)[0] == 0
    im, h, w = matching_np.padded(np.ones((200, 5), np.float32), (160, 160))
    assert (h, w) == (160, 5) and im.sum() == 800


@pytest.mark.skipif(not refload.available(), reason="/root/reference not present")
def test_oracle_equals_reference_functions():
    root = os.path.join(refload.M, "infer")
    saved = list(sys.path)
    sys.path[:0] = [os.path.join(REPO, "oracle", "refshim"), REPO, root]
    try:
        def load(name, rel):
            spec = importlib.util.spec_from_file_location(name, os.path.join(root, rel))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod
        utils, dataset = load("_ref_m_utils", "src/utils.py"), load("_ref_m_dataset", "src/dataset.py")
        query, ref, cands, len_map = make_case()
        feats, infos = utils.generate_candidates_classfiy_feature(query, ref, cands, len_map)
        ds = dataset.MatchClassifyDataset(feats, infos, (160, 160))
        want = np.stack([ds[i][0][0] for i in range(len(ds))])
        got = matching_np.classify_images(query, ref, cands, len_map, (160, 160))
        np.testing.assert_array_equal(got, want)
        assert [ds[i][1:] for i in range(len(ds))] == [(c[0], c[1]) for c in cands for _ in (0, 1)]
        meta = utils.generate_matching_feature(query, ref, len_map, cands)
        rs = dataset.MatchRefineDataset(meta, resolution=(224, 224))
        g_img, g_seg, g_h, g_w = matching_np.refine_images(query, ref, cands, len_map, (224, 224))
        for i in range(len(rs)):
            f, qid, rid, h, w = rs[i]
            np.testing.assert_array_equal(f[0], g_img[i])
            assert (qid, rid, h, w) == (cands[i][0], cands[i][1], g_h[i], g_w[i])
        assert len(set(g_seg.tolist())) > 1
    finally:
        sys.path[:] = saved
        for n in [n for n in sys.modules if n == "vsc" or n.startswith("vsc.")]:
            del sys.modules[n]
