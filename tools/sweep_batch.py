#!/usr/bin/env python
"""Encoder throughput vs the plan's internal chunk size (max_frames): smaller chunks keep the layer's
activations L2-resident, larger ones quantise better over the 74 CTA pairs."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsc22_submission_b200 import _lib  # noqa: E402
from vsc22_submission_b200.encoder import B200ViTEncoder, VIT_B16_224_GEM, random_weights  # noqa: E402

N = 4800
w = random_weights(VIT_B16_224_GEM)
x = torch.randn(N, 3, 224, 224, device="cuda").clamp_(-1, 1)
for mf in [int(a) for a in sys.argv[1:]] or [48, 64, 96, 128, 160, 192, 256]:
    enc = B200ViTEncoder(VIT_B16_224_GEM, w, max_frames=mf).cuda().eval()
    for _ in range(2):
        enc(x)
    torch.cuda.synchronize()
    _lib.prof_collect(); _lib.prof_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        enc(x)
    e1.record()
    torch.cuda.synchronize()
    _lib.prof_enable(False)
    prof = _lib.prof_collect()
    ms = e0.elapsed_time(e1) / 3
    print(f"max_frames {mf:4d}: {N / ms * 1e3:8.0f} frames/s  " +
          " ".join(f"{k}={v['ms'] / 3:.1f}" for k, v in prof.items() if v["launches"]), flush=True)
    del enc
    torch.cuda.empty_cache()
