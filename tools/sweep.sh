#!/bin/bash
# chain-length sweep (BASELINE configs[4]): bench.py at several plies-per-chain settings, one line each
for L in "$@"; do
  python bench.py --plies "$L" --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null > gpurun_out/sweep_$L.json
  python - "$L" <<'PY'
import json, sys
L = sys.argv[1]
d = json.load(open(f"gpurun_out/sweep_{L}.json"))
print("plies", L, "Mpos/s", round(d["value"]), "compress_ms", round(d["compress_ms"], 2), "decompress_ms",
      round(d["decompress_ms"], 2), "binpack_bytes", d["config"]["binpack_bytes_per_gpu"], "K1", d["roofline"]["kernel_ms"])
PY
done
