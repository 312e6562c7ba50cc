"""In-tree build of libvscb200.so (nvcc, sm_100a only).  `python -m vsc22_submission_b200.build`."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libvscb200.so")
SOURCES = ["host_util.cu", "gemm.cu", "attention.cu", "attention_tc.cu", "attention_kb.cu", "attention_ws.cu", "attention_fp32.cu", "vit_kernels.cu", "vit.cu", "swin_attention.cu", "swin_kernels.cu", "swin.cu", "sim.cu", "sim_tc.cu", "sim_tc1.cu", "sim_stream.cu", "select.cu", "merge.cu", "ensemble.cu", "pair_sims.cu",
           "sort.cu", "global_topk.cu", "tn_align.cu", "resize.cu", "jpeg.cu", "index.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "-DVSCB200_NO_FAST_MATH", *os.environ.get("VSCB200_EXTRA_NVCC_FLAGS", "").split(), "-Xptxas", "-v"]


def _stamp(src: str) -> str:
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)) + ["../../include/vscb200.h"]:
        p = os.path.join(CSRC, name)
        if os.path.isfile(p) and (name.endswith((".h", ".cuh")) or os.path.basename(p) == src):
            h.update(open(p, "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(src: str, verbose: bool) -> str:
    obj = os.path.join(BUILD, src.replace(".cu", ".o"))
    stamp_file = obj + ".stamp"
    stamp = _stamp(src)
    if os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return obj
    cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(obj + ".log", "w") as f:
        f.write(log)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{log[-4000:]}")
    if verbose:
        for line in log.splitlines():
            if "spill" in line and "0 bytes spill stores, 0 bytes spill loads" not in line:
                print(f"[build] {src}: {line.strip()}")
    open(stamp_file, "w").write(stamp)
    return obj


def build(verbose: bool = True, force: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    if force:
        for f in os.listdir(BUILD):
            os.remove(os.path.join(BUILD, f))
    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
        if verbose:
            print(f"[build] linked {LIB}")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
