set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/final_tests.log
python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' > gpurun_out/final_smoke.log 2>&1
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python tools/halfkp_time.py 100000000 > gpurun_out/final_halfkp.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v5.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_walk_runs|k_emit_chains_verify|k_candidates_scan" -c 3 -s 3 -o gpurun_out/prof_v5 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:halfkp -c 2 -s 2 -o gpurun_out/halfkp_v2 python tools/halfkp_time.py 50000000 > gpurun_out/halfkp_ncu2.log 2>&1
cat gpurun_out/final_tests.log gpurun_out/final_smoke.log gpurun_out/final_bench.json gpurun_out/final_halfkp.log
