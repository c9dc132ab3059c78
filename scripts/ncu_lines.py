#!/usr/bin/env python
"""Attribute an `ncu --set full --import-source on` capture to source lines.

    python scripts/ncu_lines.py REPORT.ncu-rep [--obj build/obj/kernels.cu.o] [--kernel shadow_kernel] [--top 40]

`ncu --page source --csv` gives per-SASS-instruction counters (instructions executed, thread
instructions, stall samples) but no line numbers; `nvdisasm -gi` of the same cubin gives the line
(and the inline chain) of every SASS instruction.  The two lists are joined by instruction offset and
summed per innermost source line and per inlined function frame.  The object must be the one the
capture ran (same source, same flags).
"""
from __future__ import annotations

import argparse
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_lines(obj: str):
    """{function symbol: {offset: [(file, line), ... innermost first]}}"""
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
    out = {}
    for cb in cubins:
        txt = subprocess.run(["nvdisasm", "-gi", "-c", cb], check=True, stdout=subprocess.PIPE, text=True).stdout
        fn, chain, fresh = None, [], False
        for ln in txt.splitlines():
            m = re.match(r"\s*\.type\s+(\S+),@function", ln)
            if m:
                fn = m.group(1)
                out.setdefault(fn, {})
                chain = []
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
            if m:
                if not fresh:
                    chain, fresh = [], True
                chain.append((os.path.basename(m.group(1)), int(m.group(2))))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m and fn:
                out[fn][int(m.group(1), 16)] = (list(chain), m.group(2).strip())
                fresh = False
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--obj", default=os.path.join(ROOT, "build", "obj", "kernels.cu.o"))
    ap.add_argument("--kernel", default="", help="substring of the kernel name (first match in the report)")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--source", default=os.path.join(ROOT, "rayhs_b200", "csrc", "kernels.cu"))
    args = ap.parse_args()

    raw = subprocess.run(["ncu", "-i", args.report, "--page", "source", "--csv"], check=True, stdout=subprocess.PIPE, text=True).stdout
    # the page is a sequence of per-kernel tables, each introduced by a "Kernel Name" row
    tables, cur = [], None
    for row in csv.reader(io.StringIO(raw)):
        if not row:
            continue
        if row[0] == "Kernel Name":
            cur = {"name": row[1], "hdr": None, "rows": []}
            tables.append(cur)
        elif cur is not None and row[0] == "Address":
            cur["hdr"] = row
        elif cur is not None and cur["hdr"]:
            cur["rows"].append(row)
    tab = next((t for t in tables if args.kernel in t["name"]), None)
    if tab is None:
        sys.exit("no kernel matching %r in %s" % (args.kernel, [t["name"] for t in tables]))
    print("kernel:", tab["name"])
    hdr = tab["hdr"]
    ix = {h: i for i, h in enumerate(hdr)}
    base = int(tab["rows"][0][0], 16)

    lines = sass_lines(args.obj)
    # entry symbol: mangled name containing the kernel's short name and template argument
    short = re.search(r"(\w+)<", tab["name"]).group(1)
    bools = re.findall(r"\(bool\)([01])", tab["name"].split("(rhd::")[0])
    targ = "I" + "".join(f"Lb{b}E" for b in bools) + "E" if bools else ""
    cands = [f for f in lines if (short + targ) in f and not f.startswith("$")]
    if not cands:
        sys.exit("kernel symbol not found in object")
    entry = min(cands, key=len)
    # device functions kept out of line follow the entry in the same section at increasing offsets
    offmap = dict(lines[entry])
    for f, m in lines.items():
        if f.startswith("$" + entry):
            offmap.update(m)

    src = open(args.source).read().splitlines() if os.path.exists(args.source) else []
    by_line = collections.defaultdict(lambda: [0, 0, 0])
    by_frame = collections.defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    missing = 0
    for r in tab["rows"]:
        off = int(r[0], 16) - base
        inst = int(r[ix["Instructions Executed"]] or 0)
        tinst = int(r[ix["Thread Instructions Executed"]] or 0)
        samp = int(r[ix["# Samples"]] or 0)
        tot[0] += inst
        tot[1] += tinst
        tot[2] += samp
        info = offmap.get(off)
        if not info or not info[0]:
            missing += inst
            continue
        chain = info[0]
        for k, v in enumerate((inst, tinst, samp)):
            by_line[chain[0]][k] += v
        for fr in set(chain):
            for k, v in enumerate((inst, tinst, samp)):
                by_frame[fr][k] += v
    print(f"total: {tot[0]:.3e} warp inst, {tot[1]:.3e} thread inst ({tot[1] / max(tot[0], 1):.1f} threads/inst), {tot[2]} samples; "
          f"unattributed {missing / max(tot[0], 1):.1%}")

    def show(title, table):
        print(f"\n== {title}: % warp-inst | % samples | threads/inst | line")
        for (f, ln), (i, t, s) in sorted(table.items(), key=lambda kv: -kv[1][2])[: args.top]:
            text = src[ln - 1].strip()[:100] if f == os.path.basename(args.source) and 0 < ln <= len(src) else ""
            print(f"{100 * i / max(tot[0], 1):6.2f} {100 * s / max(tot[2], 1):6.2f} {t / max(i, 1):5.1f}  {f}:{ln}  {text}")

    show("innermost line", by_line)
    show("inclusive (any frame of the inline chain)", by_frame)


if __name__ == "__main__":
    main()
