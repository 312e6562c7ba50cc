#!/usr/bin/env python
"""Short similarity runs for ncu captures (never a benchmark number): config 3 search + streaming form."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsc22_submission_b200 import search
g = torch.Generator(device="cuda").manual_seed(0)
unit = lambda n: torch.nn.functional.normalize(torch.randn((n, 512), generator=g, device="cuda"))
Q, R, Z = unit(10000), unit(40000), unit(40000)
q_t, r_t, _ = search.score_normalize_tensors(Q, R, Z, beta=1.2, nk=1)
ix = search.DeviceIndex(512); ix.add(r_t)
for _ in range(2):
    D, I = ix.search(q_t, 10)
print("fallbacks (score-normalised refs, k=10):", ix.last_fallbacks())
zi = search.DeviceIndex(512); zi.add(unit(40000))
zi.search(Q, 1)
print("fallbacks (unit bank, k=1):", zi.last_fallbacks())
torch.cuda.synchronize()
for name, fn in (("k=10 score-normalised", lambda: ix.search(q_t, 10)), ("k=1 unit", lambda: zi.search(Q, 1))):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    print(name, "search call: %.3f ms" % e0.elapsed_time(e1))
big = search.DeviceIndex(512); big.add(unit(1_000_000))
for _ in range(2):
    big.search(Q[:40], 10)
torch.cuda.synchronize()
print("done")
