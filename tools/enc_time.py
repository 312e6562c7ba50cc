#!/usr/bin/env python
"""ViT-B/16 encoder throughput (10k frames, chunk 256) with and without the per-kernel event profiler."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsc22_submission_b200 import _lib
from vsc22_submission_b200.encoder import B200ViTEncoder, VIT_B16_224_GEM, random_weights
N = 5120
enc = B200ViTEncoder(VIT_B16_224_GEM, random_weights(VIT_B16_224_GEM), max_frames=256).cuda().eval()
x = torch.randn(N, 3, 224, 224, device="cuda").clamp_(-1, 1)
for _ in range(3): enc(x)
for prof in (False, True, False):
    torch.cuda.synchronize(); _lib.prof_collect(); _lib.prof_enable(prof)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4): enc(x)
    e1.record(); torch.cuda.synchronize(); _lib.prof_enable(False)
    ms = e0.elapsed_time(e1) / 4
    print(f"PDL={os.environ.get('VSCB200_PDL','1')} prof={prof}: {N / ms * 1e3:.0f} frames/s ({ms:.1f} ms per {N})", flush=True)
