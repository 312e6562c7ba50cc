#!/usr/bin/env python
"""Where does the host-API (forward_host) time go?  device-resident forward vs H2D alone vs the pipelined call."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsc22_submission_b200.encoder import B200ViTEncoder, VIT_B16_224_GEM, random_weights
N = 5120
enc = B200ViTEncoder(VIT_B16_224_GEM, random_weights(VIT_B16_224_GEM), max_frames=256).cuda().eval()
xd = torch.randn(N, 3, 224, 224, device="cuda").clamp_(-1, 1)
xh = torch.empty((N, 3, 224, 224), dtype=torch.float32, pin_memory=True); xh.copy_(xd); torch.cuda.synchronize()
xn = xh.numpy()
def t(fn, reps=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
print(f"device-resident forward: {t(lambda: enc(xd)):.1f} ms")
buf = torch.empty((256, 3, 224, 224), device="cuda")
def h2d():
    for i in range(0, N, 256): buf.copy_(xh[i:i + 256], non_blocking=True)
print(f"H2D alone (20 x 154 MB pinned): {t(h2d):.1f} ms")
print(f"forward_host (pinned): {t(lambda: enc.forward_host(xn, 'cuda:0')):.1f} ms")
s2 = torch.cuda.Stream()
def both():
    with torch.cuda.stream(s2):
        h2d()
    enc(xd)
print(f"device forward + concurrent H2D on another stream: {t(both):.1f} ms")
