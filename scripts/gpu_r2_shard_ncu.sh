#!/bin/bash
# Launch list of one frame of shard 0 of 8 of the bench frame on one GPU (what a rank of the 8-GPU run executes):
# per-launch duration, issue and warp-slot utilisation — where the strong-scaling loss sits.
mkdir -p gpurun_out
python scripts/shard_frame.py --shards 8 --frames 3 > gpurun_out/shard8_plain.json 2>&1 || { tail -3 gpurun_out/shard8_plain.json; exit 1; }
tail -1 gpurun_out/shard8_plain.json
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__cycles_active.avg,sm__cycles_elapsed.max
ncu --metrics $M --clock-control none -k regex:"trace_kernel|classify_kernel|shadow_|resolve_kernel" -s 156 -c 13 --csv --log-file gpurun_out/launches_shard8_${TAG:-r2}.csv python scripts/shard_frame.py --shards 8 --frames 3 > gpurun_out/ncu_shard8.log 2>&1; echo "rc=$?"
ncu --metrics $M --clock-control none -k regex:"trace_kernel|classify_kernel|shadow_|resolve_kernel" -s 52 -c 13 --csv --log-file gpurun_out/launches_shard1_${TAG:-r2}.csv python scripts/shard_frame.py --shards 8 --frames 3 > gpurun_out/ncu_shard1.log 2>&1; echo "rc=$?"
