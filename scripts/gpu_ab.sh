#!/bin/bash
# A/B: time one C4 frame for every library variant under variants/ (and the default build)
mkdir -p gpurun_out
[ -n "$AB_TESTS" ] && timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
FMT="import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('  total %.2f trace %.2f shadow %.2f resolve %.2f' % (d['ms_total'], d['ms_trace'], d['ms_shadow'], d['ms_resolve']))"
echo "default"; python scripts/profile_frame.py --frames 5 | python -c "$FMT"
for v in variants/*/; do
  n=$(basename $v)
  echo "$n"
  RAYHS_B200_LIB=$PWD/variants/$n/librayhs_b200.so python scripts/profile_frame.py --frames 5 2>&1 | python -c "$FMT"
done
