#!/usr/bin/env python
"""Where the host-API similarity step (bench.py sim.e2e) spends its time: score_normalize on per-video host arrays,
IndexFlat.add per video, search on numpy arrays."""
import dataclasses, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsc22_submission_b200 import faiss_compat as faiss, search

@dataclasses.dataclass
class VF:
    video_id: str
    feature: np.ndarray

rng = np.random.default_rng(0)
unit = lambda n: (lambda x: x / np.linalg.norm(x, axis=1, keepdims=True))(rng.standard_normal((n, 512)).astype(np.float32))
Q, R, Z = unit(10000), unit(40000), unit(40000)
vids = lambda pre, x, per: [VF(f"{pre}{i}", x[i:i + per]) for i in range(0, x.shape[0], per)]
qv, rv, zv = vids("Q", Q, 100), vids("R", R, 400), vids("N", Z, 400)
def step(prof=None):
    t0 = time.perf_counter()
    q2, r2 = search.score_normalize(qv, rv, zv, beta=1.2, nk=1)
    t1 = time.perf_counter()
    ri = faiss.IndexFlat(512, faiss.METRIC_INNER_PRODUCT)
    for r in r2:
        ri.add(r.feature)
    t2 = time.perf_counter()
    qq = np.concatenate([q.feature for q in q2], axis=0)
    t3 = time.perf_counter()
    D, I = ri.search(qq, 10)
    t4 = time.perf_counter()
    if prof is not None:
        prof.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3))
    return D, I
for _ in range(3): step()
p = []
for _ in range(5): step(p)
a = np.array(p).mean(0) * 1e3
print("score_normalize %.2f ms | %d x IndexFlat.add %.2f ms | concatenate queries %.2f ms | search %.2f ms | total %.2f ms" % (a[0], len(rv), a[1], a[2], a[3], a.sum()))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); step(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
