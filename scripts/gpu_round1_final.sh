#!/bin/bash
# Round-1 final evidence (tag k): bench line, launch list of the same workload, full captures of the dominant kernels.
# Every ncu pass runs only after the same command has exited 0 without ncu.
TAG=${TAG:-k}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"
python scripts/profile_frame.py --frames 2 > gpurun_out/pf_${TAG}.log 2>&1 || exit 1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none -s 15 -c 15 --csv --log-file gpurun_out/launches_${TAG}.csv python scripts/profile_frame.py --frames 2 > gpurun_out/ncu_l_${TAG}.log 2>&1; echo "launch list rc=$?"
for spec in shadow_kernel_fast:7:shadow_pass0 shadow_kernel_fast:8:shadow_pass1 trace_kernel:7:trace_pass0 trace_kernel:8:trace_pass1; do
  IFS=: read k s tag <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -f -o gpurun_out/prof_${tag}_${TAG} python scripts/profile_frame.py --frames 2 > gpurun_out/ncu_${tag}_${TAG}.log 2>&1
  echo "ncu $tag rc=$?"
done
# the split schedule on the incoherent synthetic scene (1 M triangles + 1 k spheres, 1080p, 4 spp)
C5="python scripts/c5_perf.py --tris 1000000 --width 1920 --height 1080 --spp 4 --frames 2"
$C5 --shadow split > gpurun_out/c5_split_${TAG}.json 2>&1 && $C5 --shadow pooled > gpurun_out/c5_pooled_${TAG}.json 2>&1
ncu --metrics $M --clock-control none -s 29 -c 8 --csv --log-file gpurun_out/launches_c5_split_${TAG}.csv $C5 --shadow split > /dev/null 2>&1; echo "c5 split list rc=$?"
ncu --metrics $M --clock-control none -s 15 -c 4 --csv --log-file gpurun_out/launches_c5_pooled_${TAG}.csv $C5 --shadow pooled > /dev/null 2>&1; echo "c5 pooled list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:shadow_walk_kernel -s 7 -c 1 -f -o gpurun_out/prof_c5_walk_${TAG} $C5 --shadow split > /dev/null 2>&1; echo "ncu c5 walk rc=$?"
