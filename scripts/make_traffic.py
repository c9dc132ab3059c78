#!/usr/bin/env python
"""profiles/traffic.json from an ncu launch list (scripts/gpu_round1_final.sh): DRAM bytes per launch of each kernel of
one frame of the bench workload, which bench.py reports as roofline.traffic.
    python scripts/make_traffic.py gpurun_out/launches_k.csv profiles/r1k_ncu_launches.csv"""
import collections, csv, json, os, shutil, sys
src, dst = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
L = collections.OrderedDict()
for r in rows[hi + 1:]:
    d = dict(zip(hdr, r))
    name = d["Kernel Name"].split("::")[-1].split("<")[0].split("(")[0]
    L.setdefault(int(d["ID"]), {"k": name})[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
agg = collections.defaultdict(lambda: collections.Counter())
for x in L.values():
    a = agg[x["k"]]
    a["launches"] += 1
    a["dram_bytes"] += x["dram__bytes_read.sum"] + x["dram__bytes_write.sum"]
    a["l2_bytes"] += x.get("lts__t_bytes.sum", 0)
    a["ns"] += x["gpu__time_duration.sum"]
    a["warp_inst"] += x["smsp__inst_executed.sum"]
out = {"source": os.path.basename(dst), "workload": "one frame of bench.py's workload (one chunk), ncu --clock-control none", "kernels": {}}
for k, a in agg.items():
    out["kernels"][k] = {"launches": int(a["launches"]), "dram_bytes_per_launch": a["dram_bytes"] / a["launches"],
                         "l2_bytes_per_launch": a["l2_bytes"] / a["launches"], "ms_per_frame_under_ncu": a["ns"] / 1e6,
                         "warp_instructions_per_frame": a["warp_inst"]}
shutil.copy(src, dst)
with open(os.path.join(os.path.dirname(dst), "traffic.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
