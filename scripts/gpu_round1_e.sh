#!/bin/bash
# round-1 checkpoint e: GPU parity suite, bench line, full ncu captures of wall-only and dragon chunks
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; echo "bench rc=$?"
cut -c1-600 gpurun_out/bench_e.json
for spec in shadow:112:wall shadow:175:mid trace:112:wall trace:175:mid; do
  IFS=: read k s tag <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:${k}_kernel -s $s -c 1 -f -o gpurun_out/prof_${k}_${tag}_e python scripts/profile_frame.py --frames 2 > gpurun_out/ncu_${k}_${tag}_e.log 2>&1
  echo "ncu $k $tag rc=$?"
done
