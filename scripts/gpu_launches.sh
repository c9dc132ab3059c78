#!/bin/bash
mkdir -p gpurun_out
python scripts/profile_frame.py --frames 2 > gpurun_out/pf_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -s 240 -c 240 --csv --log-file gpurun_out/launches_d.csv python scripts/profile_frame.py --frames 2 > gpurun_out/ncu_l.log 2>&1
echo "launches rc=$?"
python scripts/profile_frame.py --frames 2 --count | tail -1
