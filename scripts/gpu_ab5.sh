#!/bin/bash
# GPU tests, then the bench frame (C4, pooled shadows) for the default build and every variant; light-map resolutions.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
FMT="import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('  total %.2f trace %.2f shadow %.2f resolve %.2f' % (d['ms_total'], d['ms_trace'], d['ms_shadow'], d['ms_resolve']))"
echo "default"; python scripts/profile_frame.py --frames 5 | python -c "$FMT"
echo "default res1024"; RAYHS_B200_LIGHT_MAP_RES=1024 python scripts/profile_frame.py --frames 5 | python -c "$FMT"
echo "default split"; python scripts/profile_frame.py --frames 5 --shadow split | python -c "$FMT"
for v in variants/*/; do
  n=$(basename $v)
  echo "$n"
  RAYHS_B200_LIB=$PWD/variants/$n/librayhs_b200.so python scripts/profile_frame.py --frames 5 2>&1 | python -c "$FMT"
done
