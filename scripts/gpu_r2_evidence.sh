#!/bin/bash
# Round-2 evidence set of a build on the bench workload (C4): GPU tests, smoke, the bench line, the ncu launch list of one
# frame, ncu --set full of pass 0 / pass 1 of the three traversal kernels.  Every ncu pass runs only after the same
# command has exited 0 without ncu.  TAG names the outputs.
TAG=${TAG:-r2}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py --smoke 2>&1 | tail -1
TAG=$TAG ARGS="$BENCH_ARGS" bash scripts/gpu_r2_bench.sh
python scripts/profile_frame.py --frames 3 > gpurun_out/pf_${TAG}.log 2>&1 || { tail -5 gpurun_out/pf_${TAG}.log; exit 1; }
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum,l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum,l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum
ncu --metrics $M --clock-control none -k regex:"trace_kernel|classify_kernel|shadow_|resolve_kernel" -s 26 -c 13 --csv --log-file gpurun_out/launches_${TAG}.csv python scripts/profile_frame.py --frames 3 > gpurun_out/ncu_l_${TAG}.log 2>&1; echo "launch list rc=$?"
for spec in ${SPECS:-trace_kernel:8:trace_pass0 classify_kernel:8:classify_pass0 shadow_pooled:8:pooled_pass0 trace_kernel:9:trace_pass1 shadow_pooled:9:pooled_pass1}; do
  IFS=: read k s tag <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -f -o gpurun_out/prof_${tag}_${TAG} python scripts/profile_frame.py --frames 3 > gpurun_out/ncu_${tag}_${TAG}.log 2>&1
  echo "ncu $tag rc=$?"
done
