# 8-GPU box: the default bench line at N = 8, then BASELINE configs[4]: chain lengths 1 / 8 / 64 / 400 at 125 M positions
# per GPU = 1 B positions over the 8 GPUs (one file in both directions, see bench.py)
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_n8.json 2> gpurun_out/r2_n8.err
for L in 1 8 64 400; do
  $TR --master-port $((29530 + L % 50)) bench.py --gpus 8 --steps 3 --warmup 3 --positions 125000000 --plies $L --no-e2e --no-strong \
      > gpurun_out/r2_sweep8_L$L.json 2> gpurun_out/r2_sweep8_L$L.err
done
for f in gpurun_out/r2_n8.json gpurun_out/r2_sweep8_L*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], "value", round(d["value"]), "ms", round(d["ms_per_step"], 2), "c", round(d["compress_ms"], 2), "d", round(d["decompress_ms"], 2),
          "e2e", d.get("e2e", {}).get("value"), "strong", d.get("decompress_strong", {}).get("ms"), "check", d.get("one_file_check", {}).get("identical"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
tail -3 gpurun_out/r2_n8.err
