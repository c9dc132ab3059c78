#!/bin/bash
# ncu --set full of chosen launches of one bench frame (frame 3 of profile_frame.py --frames 3): SPECS="kernel:skip:tag ..."
TAG=${TAG:-r2}
mkdir -p gpurun_out
python scripts/profile_frame.py --frames 3 > gpurun_out/pf_${TAG}.log 2>&1 || { tail -5 gpurun_out/pf_${TAG}.log; exit 1; }
for spec in $SPECS; do
  IFS=: read k s tag <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -f -o gpurun_out/prof_${tag}_${TAG} python scripts/profile_frame.py --frames 3 > gpurun_out/ncu_${tag}_${TAG}.log 2>&1
  echo "ncu $tag rc=$?"
done
