#!/usr/bin/env python
"""vit_v68 (timm ViT-B/32 @384 + GeM 1x1-conv head, sscd.py:70-95; T = 145) throughput + kernel split."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsc22_submission_b200 import _lib
from vsc22_submission_b200.encoder import B200ViTEncoder, VIT_V68, random_weights
spec = VIT_V68
enc = B200ViTEncoder(spec, random_weights(spec), max_frames=512).cuda().eval()
N = 4096
x = torch.randn(N, 3, 384, 384, device="cuda").clamp_(-1, 1)
for _ in range(2): enc(x)
torch.cuda.synchronize(); _lib.prof_collect(); _lib.prof_enable(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(2): enc(x)
e1.record(); torch.cuda.synchronize(); _lib.prof_enable(False)
prof = _lib.prof_collect(); ms = e0.elapsed_time(e1) / 2
print(f"vit_v68: {N / ms * 1e3:.0f} frames/s  {N / ms * 1e3 * spec.flops_per_frame() / 1e12:.0f} TFLOP/s  " +
      " ".join(f"{k}={v['ms'] / 2:.1f}" for k, v in prof.items() if v["launches"]))
