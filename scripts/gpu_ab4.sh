#!/bin/bash
FMT="import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(\"  total %.2f trace %.2f shadow %.2f\" % (d[\"ms_total\"], d[\"ms_trace\"], d[\"ms_shadow\"]))"
for v in default variants/*/; do
  n=$(basename $v); echo "$n"
  if [ "$n" = default ]; then unset RAYHS_B200_LIB; else export RAYHS_B200_LIB=$PWD/variants/$n/librayhs_b200.so; fi
  python scripts/profile_frame.py --frames 4 | python -c "$FMT"
  python scripts/c5_perf.py --tris 1000000 --width 1920 --height 1080 --spp 4 --frames 3 --shadow split --trace fused | python -c "$FMT"
done
