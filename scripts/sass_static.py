#!/usr/bin/env python
"""Static SASS instruction count per source line (inclusive over the inline chain) for one kernel.
    python scripts/sass_static.py KERNEL_SUBSTRING [--obj build/obj/kernels.cu.o] [--lo N --hi M]"""
import argparse, collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_lines import sass_lines, ROOT
ap = argparse.ArgumentParser(); ap.add_argument("kernel"); ap.add_argument("--obj", default=os.path.join(ROOT, "build/obj/kernels.cu.o"))
ap.add_argument("--lo", type=int, default=0); ap.add_argument("--hi", type=int, default=10**9); ap.add_argument("--ops", action="store_true")
a = ap.parse_args()
L = sass_lines(a.obj)
fn = min([f for f in L if a.kernel in f and not f.startswith("$")], key=len)
print(fn[:120], len(L[fn]), "instructions")
src = open(os.path.join(ROOT, "rayhs_b200/csrc/kernels.cu")).read().splitlines()
incl = collections.Counter(); ops = collections.defaultdict(collections.Counter)
for off, (chain, text) in L[fn].items():
    for fr in set(chain):
        if fr[0] == "kernels.cu":
            incl[fr[1]] += 1
            ops[fr[1]][text.split()[1].split(".")[0] if text.startswith("@") else text.split()[0].split(".")[0]] += 1
for ln in sorted(incl):
    if a.lo <= ln <= a.hi:
        extra = "  " + " ".join(f"{k}:{v}" for k, v in ops[ln].most_common(6)) if a.ops else ""
        print(f"{incl[ln]:5d}  {ln:5d}  {src[ln-1].strip()[:90]}{extra}")
