#!/usr/bin/env python
"""profiles/traffic.json from an ncu launch list of ONE frame of a bench workload (scripts/gpu_r2_evidence.sh): per
kernel DRAM and L2 bytes, device time, warp / thread instructions, issue-slot and warp-slot utilisation, which bench.py
reports under roofline.measured_under_ncu.
    python scripts/make_traffic.py c4 gpurun_out/launches_X.csv profiles/r2X_ncu_launches.csv"""
import collections, csv, json, os, shutil, sys
workload, src, dst = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
L = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    name = d["Kernel Name"].split("::")[-1].split("<")[0].split("(")[0]
    try:
        L.setdefault(int(d["ID"]), {"k": name})[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
    except ValueError:
        pass
agg = collections.defaultdict(lambda: collections.Counter())
for x in L.values():
    a = agg[x["k"]]
    ns = x["gpu__time_duration.sum"]
    a["launches"] += 1
    a["dram_bytes"] += x["dram__bytes_read.sum"] + x["dram__bytes_write.sum"]
    a["l2_bytes"] += x.get("lts__t_bytes.sum", 0)
    a["ns"] += ns
    a["warp_inst"] += x["smsp__inst_executed.sum"]
    a["thread_inst"] += x.get("smsp__thread_inst_executed.sum", 0)
    a["issue_x_ns"] += x.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0) * ns
    a["warps_x_ns"] += x.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0) * ns
path = os.path.join(os.path.dirname(dst), "traffic.json")
try:
    out = json.load(open(path))
    if "kernels" in out:   # round-1 layout
        out = {}
except Exception:
    out = {}
entry = {"source": os.path.basename(dst), "workload": "one frame of bench.py --workload %s (one chunk), ncu --clock-control none" % workload,
         "kernels": {}}
for k, a in agg.items():
    entry["kernels"][k] = {"launches": int(a["launches"]), "dram_bytes_per_frame": a["dram_bytes"], "l2_bytes_per_frame": a["l2_bytes"],
                           "ms_per_frame_under_ncu": a["ns"] / 1e6, "warp_instructions_per_frame": a["warp_inst"],
                           "thread_instructions_per_frame": a["thread_inst"], "issue_active_pct": a["issue_x_ns"] / max(a["ns"], 1),
                           "warps_active_pct": a["warps_x_ns"] / max(a["ns"], 1)}
out[workload] = entry
shutil.copy(src, dst)
with open(path, "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(entry, indent=1))
