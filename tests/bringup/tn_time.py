"""Timing of the device-side localisation (pair sims + top-k + temporal network) vs the oracle port per pair."""
import dataclasses
import sys
import time

import numpy as np
import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))
from vsc22_submission_b200.localization import VCSLLocalizationMaxSim  # noqa: E402


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


def main(nq=2000, nr=8000, per_q=5, d=512):
    rng = np.random.default_rng(0)
    unit = lambda x: (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)
    refs = [VF(f"R{i}", unit(rng.standard_normal((int(rng.integers(20, 80)), d))), None) for i in range(nr)]
    queries = []
    for i in range(nq):
        n = int(rng.integers(10, 60))
        f = rng.standard_normal((n, d))
        src = refs[i % nr].feature
        L = min(n, len(src), 25)
        f[:L] = src[:L] + 0.3 * rng.standard_normal((L, d)) / np.sqrt(d)
        queries.append(VF(f"Q{i}", unit(f), None))
    for v in queries + refs:
        v.timestamps = np.arange(len(v.feature), dtype=np.float32)
    cands = [Cand(f"Q{i}", f"R{(i + j * 7) % nr}") for i in range(nq) for j in range(per_q)]
    t0 = time.perf_counter()
    loc = VCSLLocalizationMaxSim(queries, refs, model_type="TN", tn_max_step=5, min_length=4, similarity_bias=0.5)
    torch.cuda.synchronize()
    print(f"upload {1e3 * (time.perf_counter() - t0):.1f} ms; {len(cands)} candidate pairs")
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        boxes, nb, sc = loc.align(cands)
        t1 = time.perf_counter()
        m = loc.localize_all(cands)
        t2 = time.perf_counter()
        print(f"align {1e3 * (t1 - t0):.1f} ms ({len(cands) / (t1 - t0):.0f} pairs/s), localize_all {1e3 * (t2 - t1):.1f} ms, "
              f"{int(nb.sum())} boxes, {len(m)} matches", flush=True)
    from oracle import tn_np
    sims = loc.similarities(cands[:200])
    t0 = time.perf_counter()
    for _, s in sims:
        tn_np.tn(s, tn_max_step=5, min_length=4)
    dt = time.perf_counter() - t0
    print(f"oracle port (1 core): {200 / dt:.0f} pairs/s")


if __name__ == "__main__":
    main()
