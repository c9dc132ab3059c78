#!/bin/bash
FMT="import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('  total %.2f chunks %d' % (d['ms_total'], d['chunks']))"
for v in default variants/*/; do
  n=$(basename $v); echo "$n"
  if [ "$n" = default ]; then unset RAYHS_B200_LIB; else export RAYHS_B200_LIB=$PWD/variants/$n/librayhs_b200.so; fi
  for c in 0 67108864 33554432 16777216 8388608; do
    python scripts/profile_frame.py --frames 4 --chunk $c --no-profile 2>&1 | python -c "$FMT"
  done
done
