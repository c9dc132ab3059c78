#!/usr/bin/env python
"""Opcode histogram (executed warp instructions) of one kernel in an ncu --set full capture.
    python scripts/ncu_opcodes.py REPORT.ncu-rep [kernel-substring]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], check=True, stdout=subprocess.PIPE, text=True).stdout
cur = None; hdr = None; rows = []
for row in csv.reader(io.StringIO(raw)):
    if not row: continue
    if row[0] == "Kernel Name":
        if rows and want in cur: break
        cur, hdr, rows = row[1], None, []
    elif row[0] == "Address": hdr = row
    elif hdr: rows.append(row)
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); tot = 0
for r in rows:
    n = int(r[ix["Instructions Executed"]] or 0)
    src = r[ix["Source"]].split()
    op = src[1] if src[0].startswith("@") else src[0]
    op = op.split(".")[0]
    ops[op] += n; tot += n
print(cur, "total warp inst", tot)
for op, n in ops.most_common(40): print(f"{100*n/tot:6.2f}  {op}")
