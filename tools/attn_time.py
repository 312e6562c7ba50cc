#!/usr/bin/env python
"""Attention kernel timing at the bench shape (256 frames x 12 heads x T=197)."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsc22_submission_b200 import _lib
n, T, H = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (256, 197, 12)))
W = H * 64
qkv = torch.randn(n * T, 3 * W, device="cuda").bfloat16()
out = torch.empty((n * T, W), dtype=torch.bfloat16, device="cuda")
p = lambda t: C.c_void_p(t.data_ptr())
for _ in range(3):
    _lib.check(_lib.lib().vscb200_attention(p(qkv), p(out), n, T, H, 64, None))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    _lib.check(_lib.lib().vscb200_attention(p(qkv), p(out), n, T, H, 64, None))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"attention {ms*1e3:.1f} us  ({4.0*n*H*T*T*64/ms/1e9:.1f} TFLOP/s)  n={n} T={T} H={H} no_ws={os.environ.get('VSCB200_ATTN_NO_WS')} no_kb={os.environ.get('VSCB200_ATTN_NO_KB')}")
