#!/bin/bash
# configs[4]'s geometry (10 M triangles + 1 000 spheres) on one GPU: the set-up stages of rh_scene_create, frame times at
# 3840x2160 x 4 spp (tiled offsets), and the ncu launch list of one frame (DRAM / L2 bytes, hit rates, lane use).
TAG=${TAG:-r2}
mkdir -p gpurun_out
C5="python scripts/c5_perf.py --tris 10000000 --width 3840 --height 2160 --spp 4"
RAYHS_B200_DEBUG=s $C5 --frames 4 > gpurun_out/c5_10m_${TAG}.json 2> gpurun_out/c5_10m_${TAG}.err; echo rc=$?
grep scene_create gpurun_out/c5_10m_${TAG}.err
cut -c1-600 gpurun_out/c5_10m_${TAG}.json
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none -k regex:"trace_kernel|classify_kernel|shadow_|resolve_kernel" -s 39 -c 13 --csv --log-file gpurun_out/launches_c5_10m_${TAG}.csv $C5 --frames 4 > /dev/null 2>&1; echo "list rc=$?"
