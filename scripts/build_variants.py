#!/usr/bin/env python
"""Builds A/B variants of librayhs_b200.so (extra -D flags) into variants/<name>/ for one-call GPU comparisons.

    python scripts/build_variants.py name1:"-DX=1 -DY=2" name2:"..."
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rayhs_b200 import build as B  # noqa: E402


def main():
    for spec in sys.argv[1:]:
        name, flags = spec.split(":", 1)
        out = os.path.join(ROOT, "variants", name)
        os.makedirs(out, exist_ok=True)
        objs = []
        for src in B.CUDA_SOURCES:
            o = os.path.join(out, src + ".o")
            cmd = [B.NVCC, *B.NVCC_FLAGS, *flags.split(), "-c", os.path.join(B.CSRC, src), "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode:
                sys.exit(r.stdout + r.stderr)
            if src == "kernels.cu":
                regs = [l.strip() for l in (r.stdout + r.stderr).splitlines() if "registers" in l or "spill" in l]
                print(name, "|", " ; ".join(x.replace("ptxas info    : ", "") for x in regs[:4]))
            objs.append(o)
        for src in B.CXX_SOURCES:
            objs.append(os.path.join(B.OBJ, src + ".o"))
        subprocess.check_call([B.NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o",
                               os.path.join(out, "librayhs_b200.so"), *objs, "-lpthread", "-ldl"])


if __name__ == "__main__":
    main()
