#!/bin/bash
for v in default variants/*/; do
  n=$(basename $v)
  if [ "$n" = default ]; then unset RAYHS_B200_LIB; else export RAYHS_B200_LIB=$PWD/variants/$n/librayhs_b200.so; fi
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$n', 'value ms %.2f' % d['ms_per_step'], 'e2e ms %.2f' % d['e2e']['ms_per_step'])"
done
