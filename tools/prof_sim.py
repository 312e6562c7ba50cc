#!/usr/bin/env python
"""Short similarity runs for ncu captures (never a benchmark number): config 3 search + streaming form."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsc22_submission_b200 import search
g = torch.Generator(device="cuda").manual_seed(0)
unit = lambda n: torch.nn.functional.normalize(torch.randn((n, 512), generator=g, device="cuda"))
Q, R = unit(10000), unit(40000)
ix = search.DeviceIndex(512); ix.add(R)
for _ in range(2):
    ix.search(Q, 10)
big = search.DeviceIndex(512); big.add(unit(1_000_000))
for _ in range(2):
    big.search(Q[:40], 10)
torch.cuda.synchronize()
print("done")
