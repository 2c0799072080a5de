#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --windows 4 --steps 1 --warmup 3 --concurrent 1 --batch 4 --no-cpu-baseline --no-kernel-timing"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn_fused_kernel -s 20 -c 1 -o gpurun_out/prof_attn64 $CMD > gpurun_out/ncu_attn64.log 2>&1; echo "ncu attn64 exit $?"
ls -la gpurun_out/*.ncu-rep
