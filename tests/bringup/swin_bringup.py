#!/usr/bin/env python
"""Swin-V2 bring-up: small configurations against the fp32 oracle, stage by stage (truncated depths)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import swin_ref
from vsc22_submission_b200.swin_encoder import B200SwinEncoder, SwinSpec, random_weights, SWINV2_B_256

def run(name, **kw):
    spec = SwinSpec(**kw); ospec = swin_ref.SwinSpec(**kw)
    w = random_weights(spec, seed=0)
    x = torch.randn(2, 3, spec.img, spec.img, generator=torch.Generator().manual_seed(1)).clamp(-1, 1)
    ref = swin_ref.forward(ospec, w, x).numpy()
    enc = B200SwinEncoder(spec, w, max_frames=2).cuda().eval()
    out = enc(x.cuda()).cpu().numpy()
    rel = np.linalg.norm(out - ref, axis=1) / np.linalg.norm(ref, axis=1)
    print(f"{name}: rel L2 {rel}  {'OK' if rel.max() < 3e-2 else 'FAIL'}", flush=True)

run("1 stage, ws=res=8 (no windows, no shift)", img=32, patch=4, embed=64, depths=(1,), heads=(2,), window=8, pretrained_windows=(0,), out_dim=32)
run("1 stage, 2 blocks, ws 8 on 16x16 (shift+mask)", img=64, patch=4, embed=64, depths=(2,), heads=(2,), window=8, pretrained_windows=(6,), out_dim=32)
run("1 stage, ws 16 on 32x32 (N=256, two M tiles, shift)", img=128, patch=4, embed=64, depths=(2,), heads=(2,), window=16, pretrained_windows=(12,), out_dim=32)
run("2 stages with merging, ws 4", img=32, patch=4, embed=64, depths=(2, 2), heads=(2, 4), window=4, pretrained_windows=(0, 3), out_dim=32)
run("golden config", img=128, patch=4, embed=64, depths=(2, 2, 2, 2), heads=(2, 4, 8, 16), window=8, pretrained_windows=(6, 6, 6, 3), out_dim=64)
if len(sys.argv) > 1:
    enc = B200SwinEncoder(SWINV2_B_256, random_weights(SWINV2_B_256), max_frames=64).cuda().eval()
    x = torch.randn(256, 3, 256, 256, device="cuda").clamp_(-1, 1)
    for _ in range(2): enc(x)
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(3): enc(x)
    torch.cuda.synchronize(); dt = (time.time() - t0) / 3
    print(f"SwinV2-B 256: {256 / dt:.0f} frames/s, {256 / dt * SWINV2_B_256.flops_per_frame() / 1e12:.1f} TFLOP/s")
