#!/bin/bash
[ -n "$AB_TESTS" ] && timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
FMT="import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('  total %.2f trace %.2f shadow %.2f | nodes %.2fG+%.2fG tris %.2fG+%.2fG' % (d['ms_total'], d['ms_trace'], d['ms_shadow'], d['node_visits']/1e9, d['shadow_node_visits']/1e9, d['tri_tests']/1e9, d['shadow_tri_tests']/1e9))"
for v in default variants/*/; do
  n=$(basename $v); echo "$n"
  if [ "$n" = default ]; then unset RAYHS_B200_LIB; else export RAYHS_B200_LIB=$PWD/variants/$n/librayhs_b200.so; fi
  python scripts/profile_frame.py --frames 4 2>&1 | python -c "$FMT"
  python scripts/profile_frame.py --frames 2 --count 2>&1 | python -c "$FMT"
  python scripts/c5_perf.py --tris 1000000 --width 1920 --height 1080 --spp 4 --frames 3 --shadow split 2>&1 | python -c "$FMT"
  python scripts/c5_perf.py --tris 1000000 --width 1920 --height 1080 --spp 4 --frames 2 --shadow split --count 2>&1 | python -c "$FMT"
done
