#!/bin/bash
# bench.py under torchrun on N GPUs of one box (gpurun --gpus N -- 'N=8 WORKLOAD=c4 TAG=x bash scripts/gpu_r2_multi.sh').
N=${N:-2}
WORKLOAD=${WORKLOAD:-c4}
TAG=${TAG:-r2}
mkdir -p gpurun_out
timeout ${BENCH_TIMEOUT:-420} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --workload $WORKLOAD $ARGS > gpurun_out/bench_${WORKLOAD}_${N}gpu_${TAG}.json 2> gpurun_out/bench_${WORKLOAD}_${N}gpu_${TAG}.err
echo "bench rc=$?"
tail -3 gpurun_out/bench_${WORKLOAD}_${N}gpu_${TAG}.err | cut -c1-300
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${WORKLOAD}_${N}gpu_${TAG}.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "parity", "exchange")})
    print("e2e", d["e2e"]["ms_per_step"], "assembled", d.get("assembled_frame_equals_single_gpu_frame"), d.get("both_exchanges_assemble_the_same_frame"))
    print("per rank", d.get("per_rank_kernel_ms"))
    print("setup", d.get("setup"))
except Exception as e:
    print("no line:", e)
PY
