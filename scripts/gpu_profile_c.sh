#!/bin/bash
mkdir -p gpurun_out
python scripts/profile_frame.py --frames 2 > gpurun_out/pf_plain.log 2>&1 && tail -1 gpurun_out/pf_plain.log | cut -c1-400 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 240 -c 240 --csv --log-file gpurun_out/launches_c.csv python scripts/profile_frame.py --frames 2 > gpurun_out/ncu_l.log 2>&1
echo "launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:shadow_kernel -s 168 -c 1 -o gpurun_out/prof_shadow_c python scripts/profile_frame.py --frames 2 > gpurun_out/ncu_shadow.log 2>&1
echo "ncu shadow rc=$?"
ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 168 -c 1 -o gpurun_out/prof_trace_c python scripts/profile_frame.py --frames 2 > gpurun_out/ncu_trace.log 2>&1
echo "ncu trace rc=$?"
