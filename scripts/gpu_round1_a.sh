#!/bin/bash
# first GPU measurement pass: tests, smoke, bench, ncu launch list, ncu full capture of both traversal kernels
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nproc
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_a.json; tail -5 gpurun_out/bench_a.err
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 960 -c 480 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'trace_kernel|shadow_kernel' -s 964 -c 6 -o gpurun_out/prof_r1a python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out
