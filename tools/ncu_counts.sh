#!/bin/bash
# GPU box: instruction / pipe / memory counters of the dominant kernels for experiment builds
N=${N:-10000000}
K=${K:-"k_walk_runs|k_walk_chains|k_emit_chains_verify"}
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_lsu.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
for v in "$@"; do
  if [ "$v" = main ]; then unset NNP_LIB; else export NNP_LIB=$PWD/variants/$v.so; fi
  python tools/kern_probe.py $N ${PLIES:-100} ${MODE:-walk} > gpurun_out/probe_$v.log 2>&1 && \
  ncu --metrics $M --clock-control none -k regex:"$K" -s ${S:-2} -c ${C:-2} --csv --log-file gpurun_out/counts_$v.csv python tools/kern_probe.py $N ${PLIES:-100} ${MODE:-walk} > gpurun_out/ncu_counts_$v.log 2>&1
  tail -1 gpurun_out/ncu_counts_$v.log
done
