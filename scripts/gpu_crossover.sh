#!/bin/bash
FMT="import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('  %8d tris: total %.2f trace %.2f shadow %.2f' % (d['tris'], d['ms_total'], d['ms_trace'], d['ms_shadow']))"
for v in default variants/nosplit; do
  n=$(basename $v); echo "$n"
  if [ "$n" = default ]; then unset RAYHS_B200_LIB; else export RAYHS_B200_LIB=$PWD/variants/$n/librayhs_b200.so; fi
  for t in 10000 30000 100000 300000; do
    python scripts/c5_perf.py --tris $t --spheres 100 --width 1920 --height 1080 --spp 4 --frames 3 2>&1 | python -c "$FMT"
  done
done
