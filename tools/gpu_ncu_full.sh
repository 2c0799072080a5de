#!/bin/bash
# one `ncu --set full` capture of a kernel (regex $1, skip $2, count $3) from a short batch-1 bench run -> gpurun_out/prof_$4
mkdir -p gpurun_out
CMD1="python bench.py --windows 4 --steps 1 --warmup 3 --concurrent 1 --batch ${5:-1} --no-cpu-baseline --no-kernel-timing"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:$1 -s $2 -c $3 -f -o gpurun_out/prof_$4 $CMD1 > gpurun_out/ncu_full_$4.log 2>&1; echo "ncu full $4 exit $?"
ncu -i gpurun_out/prof_$4.ncu-rep --page raw --csv > gpurun_out/prof_$4.raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$4.ncu-rep --page source --csv > gpurun_out/prof_$4.source.csv 2>/dev/null
ls -la gpurun_out/prof_$4*
