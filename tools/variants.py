#!/usr/bin/env python
"""A/B builds of one translation unit: tools/variants.py <file.cu> name1="-DX=1 -DY=2" name2="..." links
vsc22_submission_b200/_variants/lib_<name>.so from the regular objects with <file.cu> recompiled under the extra flags.
Select one at run time with VSCB200_LIB=<path>."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsc22_submission_b200 import build as b

def main():
    src = sys.argv[1]
    b.build(verbose=False)
    out = os.path.join(b.HERE, "_variants")
    os.makedirs(out, exist_ok=True)
    for spec in sys.argv[2:]:
        name, flags = spec.split("=", 1)
        obj = os.path.join(out, f"{src[:-3]}_{name}.o")
        r = subprocess.run([b.NVCC, *b.FLAGS, *flags.split(), "-c", os.path.join(b.CSRC, src), "-o", obj], capture_output=True, text=True)
        if r.returncode:
            raise SystemExit(r.stdout + r.stderr)
        regs = [l for l in (r.stdout + r.stderr).splitlines() if "registers" in l or "spill" in l]
        objs = [obj if s == src else os.path.join(b.BUILD, s.replace(".cu", ".o")) for s in b.SOURCES]
        lib = os.path.join(out, f"lib_{name}.so")
        subprocess.run([b.NVCC, "-shared", "-o", lib, *objs, "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a"], check=True)
        print(name, lib, *regs[-4:], sep="\n  ")

if __name__ == "__main__":
    main()
