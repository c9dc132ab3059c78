#!/bin/bash
# bench.py on the box: JSON line to gpurun_out/bench_${TAG}.json, stderr beside it.  ARGS = extra bench.py arguments.
TAG=${TAG:-r2}
mkdir -p gpurun_out
timeout ${BENCH_TIMEOUT:-1500} python bench.py $ARGS > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_${TAG}.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${TAG}.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "parity", "setup")})
    print("e2e", d["e2e"]["ms_per_step"], "cold", d.get("cold_start"))
    print("roofline", {k: d["roofline"].get(k) for k in ("kernel", "achieved", "frac", "kernel_ms_per_frame", "all_kernels")})
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("no line:", e)
PY
