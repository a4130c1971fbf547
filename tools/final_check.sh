# GPU box: the round-end sequence -- GPU suite, smoke, bench, launch list, full captures (outputs under gpurun_out/)
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/final_tests.log
python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' > gpurun_out/final_smoke.log 2>&1
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
SHORT="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-halfkp --no-plain --no-sweep --no-file --no-strong"
$SHORT > gpurun_out/plain_short.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv $SHORT > gpurun_out/ncu_launch.log 2>&1
$SHORT > gpurun_out/plain_short.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_walk_chains|k_emit_chains_verify|k_write_payload" -c 3 -s 3 -o gpurun_out/prof_r2_final $SHORT > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/final_tests.log gpurun_out/final_smoke.log; head -c 1500 gpurun_out/final_bench.json; tail -3 gpurun_out/ncu_full.log
