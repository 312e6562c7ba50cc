"""oracle/near_dup_np.py == the reference's own source lines (extract_query_feats.py:190-199), read from the reference
file and executed here (container-only), plus reference-free sanity checks."""
import os
import textwrap

import numpy as np
import pytest

from oracle import near_dup_np, refload


def videos():
    rng = np.random.default_rng(0)
    out = []
    for n in (1, 2, 17, 60, 200):
        x = rng.standard_normal((n, 96)).astype(np.float32)
        for _ in range(n // 4):                                   # near-duplicate frames (static shots)
            a, b = rng.integers(0, n, 2)
            x[a] = x[b] + rng.uniform(0.0, 0.3) * rng.standard_normal(96).astype(np.float32)
        out.append(x)
    out.append(np.repeat(rng.standard_normal((3, 96)).astype(np.float32), 5, axis=0))     # exact duplicates
    return out


def test_duplicates_collapse():
    x = videos()[-1]
    keep = near_dup_np.keep_indices(x)
    assert len(keep) == 3 and len({k // 5 for k in keep}) == 3
    assert near_dup_np.keep_indices(videos()[0]) == [0]


@pytest.mark.skipif(not refload.available(), reason="/root/reference not present")
def test_equals_the_reference_source_lines():
    path = os.path.join(refload.D, "infer/extract_query_feats.py")
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if "feat = features / np.linalg.norm(features" in l)
    end = next(i for i, l in enumerate(lines) if "to_keep_idx = [i for i in range(len(sim_mat))" in l)
    block = textwrap.dedent("\n".join(lines[start:end + 1]))
    removed_any = 0
    for x in videos():
        env = {"np": np, "features": x, "FRAME_THRESHOLD": 0.975}
        exec(block, env)
        assert near_dup_np.keep_indices(x) == list(env["to_keep_idx"])
        removed_any += len(x) - len(env["to_keep_idx"])
    assert removed_any > 20
