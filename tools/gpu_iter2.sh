#!/bin/bash
# tests + default bench, then the same bench with the k|v precompute path switched on
bash tools/gpu_iter.sh
BDE2VID_ATTN_KVPRE=1 timeout 600 python bench.py --no-cpu-baseline --no-kernel-timing > gpurun_out/bench_kvpre.json 2> gpurun_out/bench_kvpre.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_kvpre.json").read().strip().splitlines()[-1]); print("kvpre=1 value %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]))
PY
