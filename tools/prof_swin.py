#!/usr/bin/env python
"""One 128-frame SwinV2-B forward (after a warm-up one) for ncu launch lists / captures."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsc22_submission_b200.swin_encoder import B200SwinEncoder, SWINV2_B_256, random_weights
enc = B200SwinEncoder(SWINV2_B_256, random_weights(SWINV2_B_256), max_frames=128).cuda().eval()
x = torch.randn(128, 3, 256, 256, device="cuda").clamp_(-1, 1)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    enc(x)
torch.cuda.synchronize()
print("done")
