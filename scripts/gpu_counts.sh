#!/bin/bash
# Shadow-walk counters of the bench frame: maps + occluder cache, maps only, neither.
FMT="import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print({k: d[k] for k in ('shadow_node_visits','shadow_tri_tests','shadow_box_tests','shadow_prim_tests','node_visits','tri_tests','rays_shadow','rays_shadow_culled','ms_shadow','ms_trace')})"
echo cache; python scripts/profile_frame.py --frames 2 --count | python -c "$FMT"
echo nocache; RAYHS_B200_LIB=$PWD/variants/nocache/librayhs_b200.so python scripts/profile_frame.py --frames 2 --count | python -c "$FMT"
echo nomaps; RAYHS_B200_LIGHT_MAPS=0 python scripts/profile_frame.py --frames 2 --count | python -c "$FMT"
