#!/bin/bash
# launch list of one frame of the bench workload (one chunk): time, instructions, threads/inst, issue, dram
mkdir -p gpurun_out
N=${N:-29}
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -s $N -c $N --csv --log-file gpurun_out/launches_${TAG:-h}.csv python scripts/profile_frame.py --frames 2 > gpurun_out/ncu_l_${TAG:-h}.log 2>&1
echo rc=$?
