#!/bin/bash
# quick GPU check: parity tests + a short bench (no CPU baseline)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
r=d['roofline']
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'])
print('trace_ms',r['other_kernel']['trace_ms'],'shadow_ms',r['other_kernel']['shadow_ms'],'resolve',r['other_kernel']['resolve_ms'])
print('dom',r['kernel'],'achieved',r['achieved'],'frac',r['frac'],'gatherL2',r['gather_peak_l2_resident_GBps'])
print(d['rays_by_class_rank0'], d.get('rays_shadow_culled'))
PY
tail -3 gpurun_out/bench_quick.err
