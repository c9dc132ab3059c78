#!/bin/bash
# C4 then C5 under torchrun on N GPUs of one box (gpurun --gpus N -- 'N=8 TAG=x bash scripts/gpu_r2_multi_both.sh').
N=${N:-8}
TAG=${TAG:-r2}
[ -n "$SKIP_C4" ] || N=$N WORKLOAD=c4 TAG=$TAG BENCH_TIMEOUT=300 bash scripts/gpu_r2_multi.sh
[ -n "$SKIP_C5" ] || N=$N WORKLOAD=c5 TAG=$TAG BENCH_TIMEOUT=${C5_TIMEOUT:-600} bash scripts/gpu_r2_multi.sh
