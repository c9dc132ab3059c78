#!/bin/bash
# ncu --set full of one mid-frame shadow launch for the variants named in $VARIANTS
mkdir -p gpurun_out
for n in $VARIANTS; do
  export RAYHS_B200_LIB=$PWD/variants/$n/librayhs_b200.so
  python scripts/profile_frame.py --frames 2 > gpurun_out/pf_$n.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:shadow_kernel -s 168 -c 1 -o gpurun_out/prof_shadow_$n python scripts/profile_frame.py --frames 2 > gpurun_out/ncu_$n.log 2>&1
  echo "$n rc=$?"
done
