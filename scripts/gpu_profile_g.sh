#!/bin/bash
# launch list (time, instructions, threads/inst, issue) of one frame + full captures of the pass-0 shadow launches
# of chunk 0 (walls) and chunk 9 (dragon); TAG names the outputs
TAG=${TAG:-g}
mkdir -p gpurun_out
python scripts/profile_frame.py --frames 2 > gpurun_out/pf_${TAG}.log 2>&1 && tail -1 gpurun_out/pf_${TAG}.log | cut -c1-300
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -s 240 -c 240 --csv --log-file gpurun_out/launches_${TAG}.csv python scripts/profile_frame.py --frames 2 > gpurun_out/ncu_l_${TAG}.log 2>&1
echo "launches rc=$?"
for spec in ${SPECS:-shadow:112:wall shadow:175:mid}; do
  IFS=: read k s tag <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:${k}_kernel -s $s -c 1 -f -o gpurun_out/prof_${k}_${tag}_${TAG} python scripts/profile_frame.py --frames 2 > gpurun_out/ncu_${k}_${tag}_${TAG}.log 2>&1
  echo "ncu $k $tag rc=$?"
done
