#!/bin/bash
mkdir -p gpurun_out
python scripts/c5_perf.py --tris 1000000 --width 1920 --height 1080 --spp 4 --frames 2 | cut -c1-200
python scripts/c5_perf.py --tris 1000000 --spheres 0 --width 1920 --height 1080 --spp 4 --frames 2 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('nospheres', d['ms_total'], d['ms_trace'], d['ms_shadow'])"
RAYHS_B200_LIB=$PWD/variants/nofast/librayhs_b200.so python scripts/c5_perf.py --tris 1000000 --width 1920 --height 1080 --spp 4 --frames 2 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('nofast', d['ms_total'], d['ms_trace'], d['ms_shadow'])"
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum --clock-control none -s 15 -c 15 --csv --log-file gpurun_out/launches_c5.csv python scripts/c5_perf.py --tris 1000000 --width 1920 --height 1080 --spp 4 --frames 2 > gpurun_out/ncu_c5.log 2>&1
echo rc=$?
