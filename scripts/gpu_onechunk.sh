#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,sm__cycles_elapsed.avg,smsp__warps_active.avg.per_cycle_active --clock-control none -s 15 -c 15 --csv --log-file gpurun_out/launches_onechunk.csv python scripts/profile_frame.py --frames 2 --chunk 134217728 > gpurun_out/ncu_onechunk.log 2>&1
echo rc=$?
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__cycles_active.avg,sm__cycles_elapsed.avg,smsp__warps_active.avg.per_cycle_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -s 240 -c 30 --csv --log-file gpurun_out/launches_16chunk_head.csv python scripts/profile_frame.py --frames 2 > gpurun_out/ncu_16chunk.log 2>&1
echo rc=$?
