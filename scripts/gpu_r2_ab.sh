#!/bin/bash
# Round 2 A/B: (optionally) the GPU tests, then one timed C4 frame (kernel times from per-launch events) for the default
# build and every variant under variants/.
mkdir -p gpurun_out
if [ -n "$AB_TESTS" ]; then timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/tests_${TAG:-ab}.log; fi
FMT="import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('  total %.2f trace %.2f shadow %.2f resolve %.2f  queued %d of %d hits, pairs %d' % (d['ms_total'], d['ms_trace'], d['ms_shadow'], d['ms_resolve'], d['shadow_tasks_queued'], d['shadow_tasks'], d['shadow_walk_pairs']))"
echo "default"; timeout 300 python scripts/profile_frame.py --frames 5 2>&1 | python -c "$FMT"
for v in variants/*/; do
  n=$(basename $v)
  echo "$n"
  RAYHS_B200_LIB=$PWD/variants/$n/librayhs_b200.so timeout 300 python scripts/profile_frame.py --frames 5 2>&1 | python -c "$FMT"
done
