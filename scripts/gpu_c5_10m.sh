#!/bin/bash
# configs[4] geometry (10 M triangles + 1 k spheres) on one GPU at 3840x2160, 4 spp: frame time, then an ncu launch
# list with DRAM / L2 bytes for the HBM-resident gather evidence
mkdir -p gpurun_out
C5="python scripts/c5_perf.py --tris 10000000 --width 3840 --height 2160 --spp 4"
$C5 --frames 4 > gpurun_out/c5_10m.json 2> gpurun_out/c5_10m.err; echo rc=$?; cut -c1-400 gpurun_out/c5_10m.json
$C5 --frames 2 --shadow split --count > gpurun_out/c5_10m_count.json 2>/dev/null
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none -s 29 -c 16 --csv --log-file gpurun_out/launches_c5_10m.csv $C5 --frames 2 --shadow split > /dev/null 2>&1; echo "list rc=$?"
