#!/bin/bash
FMT="import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('  total %.2f trace %.2f shadow %.2f resolve %.2f chunks %d' % (d['ms_total'], d['ms_trace'], d['ms_shadow'], d['ms_resolve'], d['chunks']))"
for c in 0 4194304 16777216 33554432 67108864 134217728; do
  echo "chunk $c"; python scripts/profile_frame.py --frames 4 --chunk $c | python -c "$FMT"
done
