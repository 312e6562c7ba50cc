#!/usr/bin/env python
"""BASELINE configs[3] encoders on one GPU: SwinV2-L/w24 @ 384 and ViT-L/16 @ 384 (T = 577) throughput with the per-kernel-kind
CUDA-event breakdown.  VSCB200_ATTN_NO_KB=1 selects the previous attention paths (fp32 FMA windows / mma.sync) for comparison."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsc22_submission_b200 import _lib
from vsc22_submission_b200 import encoder, swin_encoder

def run(name, enc, x, flops):
    for _ in range(2): enc(x)
    torch.cuda.synchronize(); _lib.prof_collect(); _lib.prof_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2): enc(x)
    e1.record(); torch.cuda.synchronize(); _lib.prof_enable(False)
    prof = _lib.prof_collect(); ms = e0.elapsed_time(e1) / 2
    n = x.shape[0]
    print(f"{name}: {n / ms * 1e3:.0f} frames/s  {n / ms * 1e3 * flops / 1e12:.0f} TFLOP/s  ms per {n} frames {ms:.1f}  " +
          " ".join(f"{k}={v['ms'] / 2:.1f}" for k, v in prof.items() if v["launches"]), flush=True)

which = sys.argv[1] if len(sys.argv) > 1 else "both"
kb = "kb off" if os.environ.get("VSCB200_ATTN_NO_KB") else "kb on"
if which in ("both", "swin"):
    sp = swin_encoder.SWINV2_L_384
    enc = swin_encoder.B200SwinEncoder(sp, swin_encoder.random_weights(sp), max_frames=64).cuda().eval()
    run(f"SwinV2-L/w24@384 ({kb})", enc, torch.randn(256, 3, 384, 384, device="cuda").clamp_(-1, 1), sp.flops_per_frame())
    del enc
if which in ("both", "vit"):
    sp = encoder.VIT_L16_384
    enc = encoder.B200ViTEncoder(sp, encoder.random_weights(sp), max_frames=64).cuda().eval()
    run(f"ViT-L/16@384 ({kb})", enc, torch.randn(256, 3, 384, 384, device="cuda").clamp_(-1, 1), sp.flops_per_frame())
