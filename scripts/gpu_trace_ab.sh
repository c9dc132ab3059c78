#!/bin/bash
FMT="import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('  total %.2f trace %.2f shadow %.2f (trace_split %d shadow_split %d)' % (d['ms_total'], d['ms_trace'], d['ms_shadow'], d['trace_split'], d['shadow_split']))"
for t in fused split; do
  echo "dragon trace=$t"; python scripts/profile_frame.py --frames 4 --shadow pooled --trace $t | python -c "$FMT"
  echo "synthetic 1M trace=$t"; python scripts/c5_perf.py --tris 1000000 --width 1920 --height 1080 --spp 4 --frames 3 --shadow split --trace $t | python -c "$FMT"
done
echo "auto dragon"; python scripts/profile_frame.py --frames 5 --shadow auto | python -c "$FMT"
echo "auto synthetic 10M"; python scripts/c5_perf.py --frames 4 | python -c "$FMT"
