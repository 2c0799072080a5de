#!/bin/bash
# Warm-cache launch list (ncu --cache-control none): realistic serialized per-kernel times.
# usage: gpu_launchlist.sh <windows> <batch> <out-name>
mkdir -p gpurun_out
CMD="python bench.py --windows ${1:-6} --steps 1 --warmup 3 --concurrent 1 --batch ${2:-1} --no-cpu-baseline --no-kernel-timing"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 8000 --csv --log-file gpurun_out/${3:-launches_warm}.csv $CMD > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"
