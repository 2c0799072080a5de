#!/bin/bash
mkdir -p gpurun_out
export BDE2VID_PLAN_CACHE_GB=150
for cfg in "2 6" "3 6" "2 8" "4 4" "3 8" "3 4"; do
  set -- $cfg
  timeout 400 python bench.py --steps 3 --warmup 3 --concurrent $1 --batch $2 --no-cpu-baseline --no-kernel-timing --no-single > gpurun_out/sweep_$1x$2.json 2> gpurun_out/sweep_$1x$2.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/sweep_$1x$2.json").read().strip().splitlines()[-1])
    print("sweep $1x$2 value %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]))
except Exception as e:
    print("sweep $1x$2 ERR", e)
PY
done
