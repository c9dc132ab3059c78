#!/bin/bash
# Per-launch durations of the bench frame with the per-lane-refill shadow kernel forced (N = 1 frame), to compare pass by
# pass with the pooled kernel's launch list.
mkdir -p gpurun_out
python scripts/profile_frame.py --frames 3 --shadow split > gpurun_out/pf_split.log 2>&1 || { tail -3 gpurun_out/pf_split.log; exit 1; }
tail -1 gpurun_out/pf_split.log | cut -c1-300
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none -s 26 -c 13 --csv --log-file gpurun_out/launches_split_${TAG:-r2}.csv python scripts/profile_frame.py --frames 3 --shadow split > gpurun_out/ncu_split.log 2>&1; echo "rc=$?"
