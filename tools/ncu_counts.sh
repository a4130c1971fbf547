#!/bin/bash
# GPU box: instruction / pipe counters of k_walk_runs and k_emit_chains_verify for experiment builds
N=${N:-10000000}
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_adu.sum,sm__inst_executed_pipe_cbu.sum,sm__inst_executed_pipe_uniform.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active
for v in "$@"; do
  if [ "$v" = main ]; then unset NNP_LIB; else export NNP_LIB=$PWD/variants/$v.so; fi
  python tools/kern_probe.py $N > gpurun_out/probe_$v.log 2>&1 && \
  ncu --metrics $M --clock-control none -k regex:"k_walk_runs|k_emit_chains_verify" -s 2 -c 2 --csv --log-file gpurun_out/counts_$v.csv python tools/kern_probe.py $N > gpurun_out/ncu_counts_$v.log 2>&1
  tail -2 gpurun_out/ncu_counts_$v.log
done
