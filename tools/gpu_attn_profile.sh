#!/bin/bash
# (1) streams x batch sweep of the bench; (2) full ncu captures (with source) of the two fused attention kernels and the L1 MLP kernel.
mkdir -p gpurun_out
for cfg in "2 4" "3 4" "4 2" "2 8" "4 4" "1 8"; do
  set -- $cfg
  timeout 300 python bench.py --steps 3 --warmup 3 --concurrent $1 --batch $2 --no-cpu-baseline --no-kernel-timing > gpurun_out/sweep_$1x$2.json 2> gpurun_out/sweep_$1x$2.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/sweep_$1x$2.json").read().strip().splitlines()[-1])
    print("sweep $1x$2 value %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]))
except Exception as e:
    print("sweep $1x$2 ERR", e)
PY
done
CMD="python bench.py --windows 4 --steps 1 --warmup 3 --concurrent 1 --batch 4 --no-cpu-baseline --no-kernel-timing"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn_fused_kernel -s 40 -c 4 -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn exit $?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:mlp_fused_kernel -s 16 -c 1 -o gpurun_out/prof_mlp $CMD > gpurun_out/ncu_mlp.log 2>&1; echo "ncu mlp exit $?"
ls -la gpurun_out/*.ncu-rep
