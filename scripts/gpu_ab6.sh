#!/bin/bash
# GPU tests, then the bench frame (C4) with lit-triangle flags + light maps, maps only, neither; counters of the first two.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
FMT="import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('  total %.2f trace %.2f shadow %.2f resolve %.2f' % (d['ms_total'], d['ms_trace'], d['ms_shadow'], d['ms_resolve']))"
CNT="import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print({k: d[k] for k in ('shadow_node_visits','shadow_tri_tests','shadow_box_tests')})"
for e in "" "RAYHS_B200_LIT_TRIANGLES=0" "RAYHS_B200_LIGHT_MAPS=0"; do
  echo "pooled [$e]"; env $e python scripts/profile_frame.py --frames 5 | python -c "$FMT"
done
for e in "" "RAYHS_B200_LIT_TRIANGLES=0"; do
  echo "split [$e]"; env $e python scripts/profile_frame.py --frames 5 --shadow split | python -c "$FMT"
  echo "counts [$e]"; env $e python scripts/profile_frame.py --frames 2 --count | python -c "$CNT"
done
