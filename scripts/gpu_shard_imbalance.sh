#!/bin/bash
# Per-SMSP imbalance of the launches of shard 0 of 8 of the bench frame: min / avg / max over the SMSPs of active cycles and
# executed instructions (work imbalance shows as inst max >> avg; latency tails as cycles max >> avg with inst balanced).
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__cycles_active.avg,smsp__cycles_active.max,smsp__cycles_active.min,smsp__inst_executed.avg,smsp__inst_executed.max,smsp__inst_executed.min,sm__cycles_elapsed.max,sm__cycles_active.avg,sm__cycles_active.max,sm__cycles_active.min
ncu --metrics $M --clock-control none -k regex:"trace_kernel|classify_kernel|shadow_|resolve_kernel" -s 156 -c 13 --csv --log-file gpurun_out/imbalance_shard8_${TAG:-r2}.csv python scripts/shard_frame.py --shards 8 --frames 3 > gpurun_out/ncu_imb.log 2>&1; echo "rc=$?"
