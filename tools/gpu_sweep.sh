#!/bin/bash
mkdir -p gpurun_out
for cfg in "2 4" "2 8" "1 8" "3 4" "2 6" "1 16"; do
  set -- $cfg
  timeout 400 python bench.py --steps 3 --warmup 3 --concurrent $1 --batch $2 --no-cpu-baseline --no-kernel-timing > gpurun_out/sweep_$1x$2.json 2> gpurun_out/sweep_$1x$2.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/sweep_$1x$2.json").read().strip().splitlines()[-1])
    print("sweep $1x$2 value %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]))
except Exception as e:
    print("sweep $1x$2 ERR", e)
PY
done
