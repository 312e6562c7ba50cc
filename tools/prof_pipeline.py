#!/usr/bin/env python
"""Short runs of the candidate-generation / localisation / ingest kernels for ncu captures (never a benchmark number)."""
import dataclasses
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsc22_submission_b200 import ingest, search  # noqa: E402
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


g = torch.Generator(device="cuda").manual_seed(0)
unit = lambda n: torch.nn.functional.normalize(torch.randn((n, 512), generator=g, device="cuda"))
Q, R = unit(10000), unit(40000)
ix = search.DeviceIndex(512)
ix.add(R)
for _ in range(2):
    ix.global_search(Q, 300000)
    ix.global_video_pairs(torch.arange(0, 10001, 40), torch.arange(0, 40001, 50))

rng = np.random.default_rng(0)
u = lambda x: (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)
refs = [VF(f"R{i}", u(rng.standard_normal((int(rng.integers(20, 80)), 512))), None) for i in range(2000)]
qs = []
for i in range(1000):
    f = rng.standard_normal((int(rng.integers(10, 60)), 512))
    L = min(len(f), len(refs[i].feature), 25)
    f[:L] = refs[i].feature[:L] + 0.3 * rng.standard_normal((L, 512)) / np.sqrt(512)
    qs.append(VF(f"Q{i}", u(f), None))
for v in qs + refs:
    v.timestamps = np.arange(len(v.feature), dtype=np.float32)
cands = [Cand(f"Q{i}", f"R{(i + 7 * j) % 2000}") for i in range(1000) for j in range(5)]
loc = VCSLLocalizationMaxSim(qs, refs, model_type="TN", tn_max_step=5, min_length=4, similarity_bias=0.5)
for _ in range(2):
    loc.align(cands)

frames = torch.randint(0, 256, (256, 360, 640, 3), dtype=torch.uint8, device="cuda")
pre = ingest.sscd_transform(224, 224)
for _ in range(2):
    pre(frames)
torch.cuda.synchronize()
print("done")
