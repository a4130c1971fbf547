set -x
python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -40 > gpurun_out/r2a_tests.log
tail -40 gpurun_out/r2a_tests.log
