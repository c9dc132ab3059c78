#!/bin/bash
# Round-1 evidence for builds whose frames have D + 1 passes on scenes without Transparent materials (the bench scene:
# 4 trace + 4 shadow + 1 resolve = 9 launches per frame): tests, bench line, launch list of the same workload, full
# captures of the dominant kernels.  Every ncu pass runs only after the same command has exited 0 without ncu.
TAG=${TAG:-q}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"
python scripts/profile_frame.py --frames 2 > gpurun_out/pf_${TAG}.log 2>&1 || exit 1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none -s 9 -c 9 --csv --log-file gpurun_out/launches_${TAG}.csv python scripts/profile_frame.py --frames 2 > gpurun_out/ncu_l_${TAG}.log 2>&1; echo "launch list rc=$?"
for spec in shadow_kernel_fast:4:shadow_pass0 shadow_kernel_fast:5:shadow_pass1 trace_kernel:4:trace_pass0 trace_kernel:5:trace_pass1; do
  IFS=: read k s tag <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -f -o gpurun_out/prof_${tag}_${TAG} python scripts/profile_frame.py --frames 2 > gpurun_out/ncu_${tag}_${TAG}.log 2>&1
  echo "ncu $tag rc=$?"
done
