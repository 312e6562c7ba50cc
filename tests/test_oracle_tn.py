"""oracle/tn_np.py == the reference's vcsl.vta.tn (networkx) on random, planted and degenerate similarity matrices
(container-only: needs /root/reference), plus reference-free structural checks."""
import numpy as np
import pytest

from oracle import refload, tn_np

needs_ref = pytest.mark.skipif(not refload.available(), reason="/root/reference not present")


def _cases():
    rng = np.random.default_rng(0)
    out = []
    for t in range(60):
        Q, R = int(rng.integers(1, 60)), int(rng.integers(1, 70))
        s = rng.uniform(-0.2, 0.6, (Q, R)).astype(np.float32)
        for _ in range(int(rng.integers(0, 3))):               # planted diagonal segments (copied clips)
            L = int(rng.integers(3, 25))
            q0, r0 = int(rng.integers(0, max(Q - L, 1))), int(rng.integers(0, max(R - L, 1)))
            for i in range(min(L, Q - q0, R - r0)):
                s[q0 + i, r0 + i] = rng.uniform(0.6, 1.0)
        out.append(s)
    out.append(np.zeros((10, 10), np.float32))                  # nothing passes min_sim
    out.append(np.full((12, 3), 0.9, np.float32))               # fewer reference frames than top-k, all ties
    out.append(np.eye(30, dtype=np.float32))                    # a clean diagonal
    out.append(np.tile(np.eye(20, dtype=np.float32), (1, 2)) * 0.8 + 0.1)   # two parallel diagonals
    return out


PARAMS = [dict(tn_max_step=5, min_length=4), dict(), dict(tn_max_step=3, tn_top_k=2, min_length=1, max_path=3),
          dict(tn_max_step=10, tn_top_k=5, min_sim=0.5, max_iou=0.9)]


@needs_ref
def test_tn_equals_reference_function():
    refload.vsc_package("D_infer")
    try:
        from vcsl.vta import tn as ref_tn
        n = 0
        for s in _cases():
            for kw in PARAMS:
                with np.errstate(all="ignore"):
                    ref = ref_tn(s, **kw)
                assert tn_np.tn(s, **kw) == ref, (s.shape, kw)
                n += len(ref)
        assert n > 50          # the cases do produce boxes
    finally:
        refload.unload_vsc()


def test_clean_diagonal_is_found():
    s = np.full((40, 50), 0.05, np.float32)
    for i in range(25):
        s[5 + i, 10 + i] = 0.9
    boxes = tn_np.tn(s, tn_max_step=5, min_length=4)
    assert boxes and boxes[0] == [5, 10, 29, 34]
    assert tn_np.tn(np.zeros((8, 8), np.float32)) == []
    assert tn_np.tn(np.full((1, 1), 0.9, np.float32)) == []
