"""Global candidate search at the config-3 shape: per-stage profile (library ProfScope) and wall time."""
import sys, time, torch
sys.path.insert(0, ".")
from vsc22_submission_b200 import _lib
from vsc22_submission_b200.search import DeviceIndex
g = torch.Generator(device="cuda").manual_seed(2)
q = torch.nn.functional.normalize(torch.randn(10000, 512, device="cuda", generator=g))
r = torch.nn.functional.normalize(torch.randn(40000, 512, device="cuda", generator=g))
ix = DeviceIndex(512, 0); ix.add(r)
for i in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s, qi, ri = ix.global_search(q, 300000)
    torch.cuda.synchronize(); print(f"global_search {1e3*(time.perf_counter()-t0):.2f} ms n={s.numel()}", flush=True)
_lib.prof_collect(); _lib.prof_enable(True)
ix.global_search(q, 300000); torch.cuda.synchronize()
_lib.prof_enable(False); print(_lib.prof_collect())
