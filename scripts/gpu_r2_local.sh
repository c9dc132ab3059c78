#!/bin/bash
# Round 2, "before" evidence for VERDICT item 3: where do trace pass 0's DRAM writes come from?  Local-memory (stack +
# spill) traffic against global traffic per launch of one bench frame.  TAG names the outputs.
TAG=${TAG:-r2a}
mkdir -p gpurun_out
python scripts/profile_frame.py --frames 3 > gpurun_out/pf_${TAG}.log 2>&1 || { tail -5 gpurun_out/pf_${TAG}.log; exit 1; }
tail -1 gpurun_out/pf_${TAG}.log | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('frame %.2f trace %.2f shadow %.2f resolve %.2f' % (d['ms_total'], d['ms_trace'], d['ms_shadow'], d['ms_resolve']))"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum,l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum,l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum,l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sectors_srcunit_tex_lookup_miss.sum,lts__t_sectors_srcunit_tex_lookup_hit.sum
ncu --metrics $M --clock-control none -s 18 -c 9 --csv --log-file gpurun_out/launches_local_${TAG}.csv python scripts/profile_frame.py --frames 3 > gpurun_out/ncu_local_${TAG}.log 2>&1; echo "launch list rc=$?"
