#!/bin/bash
# usage: gpu_prof_one.sh <kernel regex> <launch skip> <out name>   -- one ncu --set full capture (with source) of the bench step
mkdir -p gpurun_out
CMD="python bench.py --windows 4 --steps 1 --warmup 3 --concurrent 1 --batch 4 --no-cpu-baseline --no-kernel-timing"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:$1 -s $2 -c 1 -o gpurun_out/$3 $CMD > gpurun_out/ncu_$3.log 2>&1; echo "ncu $3 exit $?"
ncu -i gpurun_out/$3.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/$3.src.csv 2>/dev/null
ncu -i gpurun_out/$3.ncu-rep --page raw --csv > gpurun_out/$3.raw.csv 2>/dev/null
ls -la gpurun_out/$3.*
