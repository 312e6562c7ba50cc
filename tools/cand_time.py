"""Timing of the device-side global candidate search (csrc/global_topk.cu) at BASELINE config-3 shapes and larger."""
import sys
import time

import torch

sys.path.insert(0, ".")
from vsc22_submission_b200.search import DeviceIndex  # noqa: E402


def run(nq, nr, K, frames_q=40, frames_r=50, reps=3):
    g = torch.Generator(device="cuda").manual_seed(2)
    q = torch.nn.functional.normalize(torch.randn(nq, 512, device="cuda", generator=g))
    r = torch.nn.functional.normalize(torch.randn(nr, 512, device="cuda", generator=g))
    ix = DeviceIndex(512, 0)
    ix.add(r)
    qo, ro = torch.arange(0, nq + 1, frames_q), torch.arange(0, nr + 1, frames_r)
    for i in range(reps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        s, qi, ri = ix.global_search(q, K)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        sc, qv, rv = ix.global_video_pairs(qo, ro)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        if i:
            print(f"nq={nq} nr={nr} K={K}: global_search {1e3 * (t1 - t0):.2f} ms ({nq * nr / (t1 - t0) / 1e9:.1f} Gpairs/s), "
                  f"video_pairs {1e3 * (t2 - t1):.2f} ms -> {sc.numel()} candidates", flush=True)
    t0 = time.perf_counter()
    S = ix.scores(q[: min(nq, 10000)])
    torch.cuda.synchronize()
    print(f"   dense scores of {min(nq, 10000)} rows alone: {1e3 * (time.perf_counter() - t0):.2f} ms")


if __name__ == "__main__":
    run(10000, 40000, 300000)
    run(40000, 400000, 1200000)
    run(40000, 400000, 12000000, reps=2)
