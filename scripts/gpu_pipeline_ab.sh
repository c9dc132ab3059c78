#!/bin/bash
# Passes pipelined over two streams (default) against one stream per chunk (RAYHS_B200_PIPELINE=0): the full bench frame,
# shard 0 of 8 of it, and the streamed (e2e) frame.
[ -n "$AB_TESTS" ] && timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
FMT="import sys,json; d=json.loads(sys.stdin.readlines()[-1]); a=d['shards_1']; b=d['shards_8']; print('  full %.2f   shard8 %.3f   eff %.3f' % (a['device_ms_per_call'], b['device_ms_per_call'], d['efficiency_device']))"
for p in 1 0; do
  echo "pipeline $p"
  RAYHS_B200_PIPELINE=$p python scripts/shard_frame.py --frames 10 2>/dev/null | python -c "$FMT"
  RAYHS_B200_PIPELINE=$p python scripts/e2e_frame.py --frames 10 2>&1 | tail -1 | cut -c1-120
done
