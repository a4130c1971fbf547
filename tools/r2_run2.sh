set -x
python -m pytest tests/test_gpu_files.py tests/test_gpu_cli.py tests/test_gpu_sharded.py -q -x 2>&1 | tail -15
