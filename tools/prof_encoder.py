#!/usr/bin/env python
"""Short encoder / similarity run for ncu captures (never a benchmark number)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsc22_submission_b200 import search  # noqa: E402
from vsc22_submission_b200.encoder import B200ViTEncoder, VIT_B16_224_GEM, random_weights  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "encoder"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 256      # plan chunk (bench.py uses 512)
if what == "encoder":
    enc = B200ViTEncoder(VIT_B16_224_GEM, random_weights(VIT_B16_224_GEM), max_frames=chunk).cuda().eval()
    x = torch.randn(chunk, 3, 224, 224, device="cuda").clamp_(-1, 1)
    for _ in range(iters):
        enc(x)
else:
    g = torch.Generator(device="cuda").manual_seed(0)
    unit = lambda n: torch.nn.functional.normalize(torch.randn((n, 512), generator=g, device="cuda"))
    Q, R = unit(10000), unit(40000)
    ix = search.DeviceIndex(512)
    ix.add(R)
    for _ in range(iters):
        ix.search(Q, 10)
    ix.search(Q[:40], 1)
torch.cuda.synchronize()
print("done")
