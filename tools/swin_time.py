#!/usr/bin/env python
"""SwinV2-B@256 throughput with the per-kernel-kind CUDA-event breakdown."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsc22_submission_b200 import _lib
from vsc22_submission_b200.swin_encoder import B200SwinEncoder, SWINV2_B_256, random_weights
mf = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = 1024
enc = B200SwinEncoder(SWINV2_B_256, random_weights(SWINV2_B_256), max_frames=mf).cuda().eval()
x = torch.randn(N, 3, 256, 256, device="cuda").clamp_(-1, 1)
for _ in range(2): enc(x)
torch.cuda.synchronize(); _lib.prof_collect(); _lib.prof_enable(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(2): enc(x)
e1.record(); torch.cuda.synchronize(); _lib.prof_enable(False)
prof = _lib.prof_collect(); ms = e0.elapsed_time(e1) / 2
print(f"max_frames {mf}: {N / ms * 1e3:.0f} frames/s  {N / ms * 1e3 * SWINV2_B_256.flops_per_frame() / 1e12:.0f} TFLOP/s  ms/1024 frames {ms:.1f}  " +
      " ".join(f"{k}={v['ms'] / 2:.1f}" for k, v in prof.items() if v["launches"]) +
      f"  gemm TF {prof['gemm']['work'] / prof['gemm']['ms'] / 1e9:.0f}")
