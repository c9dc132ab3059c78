#!/bin/bash
# Light maps A/B on the bench frame (C4) and the 1 M-triangle synthetic scene: maps on (three resolutions) and off.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "light_maps" -s 2>&1 | grep -E "shadow node visits|passed|failed|Error|assert" | tail -12
FMT="import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(\"  total %.2f trace %.2f shadow %.2f resolve %.2f\" % (d[\"ms_total\"], d[\"ms_trace\"], d[\"ms_shadow\"], d[\"ms_resolve\"]))"
for cfg in "on512:" "off:RAYHS_B200_LIGHT_MAPS=0" "on256:RAYHS_B200_LIGHT_MAP_RES=256" "on1024:RAYHS_B200_LIGHT_MAP_RES=1024"; do
  n=${cfg%%:*}; e=${cfg#*:}; echo "$n"
  env $e python scripts/profile_frame.py --frames 4 | python -c "$FMT"
  env $e python scripts/profile_frame.py --frames 4 --shadow split | python -c "$FMT"
done
for e in "" "RAYHS_B200_LIGHT_MAPS=0"; do
  echo "c5 1M $e"
  env $e python scripts/c5_perf.py --tris 1000000 --width 1920 --height 1080 --spp 4 --frames 3 --shadow split --trace fused | python -c "$FMT"
done
