#!/bin/bash
mkdir -p gpurun_out
python scripts/profile_frame.py --frames 2 > gpurun_out/pf_plain.log 2>&1 && tail -1 gpurun_out/pf_plain.log &&
ncu --set full --clock-control none --import-source on -k regex:shadow_kernel -s 168 -c 1 -o gpurun_out/prof_shadow python scripts/profile_frame.py --frames 2 > gpurun_out/ncu_shadow.log 2>&1
echo "ncu shadow rc=$?"
ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 168 -c 2 -o gpurun_out/prof_trace python scripts/profile_frame.py --frames 2 > gpurun_out/ncu_trace.log 2>&1
echo "ncu trace rc=$?"
python scripts/profile_frame.py --frames 2 --count | tail -1
