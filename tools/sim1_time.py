#!/usr/bin/env python
"""Times the config-3 search calls (10k x 40k x 512; k = 10 on the score-normalised bank, k = 1 on a unit bank) and the
scores kernel inside them (library profiler).  Used for A/B runs: VSCB200_LIB=<variant .so> python tools/sim1_time.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsc22_submission_b200 import search, _lib
g = torch.Generator(device="cuda").manual_seed(0)
unit = lambda n: torch.nn.functional.normalize(torch.randn((n, 512), generator=g, device="cuda"))
Q, R, Z = unit(10000), unit(40000), unit(40000)
q_t, r_t, _ = search.score_normalize_tensors(Q, R, Z, beta=1.2, nk=1)
ix = search.DeviceIndex(512); ix.add(r_t)
zi = search.DeviceIndex(512); zi.add(Z)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = []
for name, fn in (("k=10", lambda: ix.search(q_t, 10)), ("k=1", lambda: zi.search(Q, 1))):
    for _ in range(3): fn()
    ts = []
    for _ in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    out.append("%s call median %.3f ms min %.3f" % (name, ts[len(ts) // 2], ts[0]))
print(os.environ.get("VSCB200_LIB", "default").split("/")[-1], " | ".join(out), "| fallbacks", ix.last_fallbacks(), zi.last_fallbacks())
