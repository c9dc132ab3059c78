#!/bin/bash
# A/B on the full bench frame AND on shard 0 of 8 of it (what a rank of the 8-GPU run executes), for the default build and
# every library variant under variants/ (VARIANTS="a b" picks some): device ms per call, trace / shadow split.
mkdir -p gpurun_out
[ -n "$AB_TESTS" ] && timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
FMT="import sys,json; d=json.loads(sys.stdin.readlines()[-1]); a=d['shards_1']; b=d['shards_8']; print('  full %.2f (trace %.2f shadow %.2f)   shard8 %.3f (trace %.3f shadow %.3f)  eff %.3f' % (a['device_ms_per_call'], a['trace'], a['shadow'], b['device_ms_per_call'], b['trace'], b['shadow'], d['efficiency_device']))"
echo "default"; python scripts/shard_frame.py --frames 10 2>/dev/null | python -c "$FMT"
[ -n "$AB_E2E" ] && python scripts/e2e_frame.py --frames 10 2>&1 | tail -1 | cut -c1-90
for n in ${VARIANTS:-$(ls variants)}; do
  [ "$n" = wt ] && continue
  echo "$n"
  RAYHS_B200_LIB=$PWD/variants/$n/librayhs_b200.so python scripts/shard_frame.py --frames 10 2>/dev/null | python -c "$FMT"
  [ -n "$AB_E2E" ] && RAYHS_B200_LIB=$PWD/variants/$n/librayhs_b200.so python scripts/e2e_frame.py --frames 10 2>&1 | tail -1 | cut -c1-90
done
if [ -d variants/wt ]; then
  RAYHS_B200_LIB=$PWD/variants/wt/librayhs_b200.so RAYHS_B200_DEBUG=1 python scripts/shard_frame.py --frames 1 2>&1 | grep "pooled pass" | tail -8 | cut -c1-420
fi
