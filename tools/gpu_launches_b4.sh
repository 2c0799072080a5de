#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --windows 6 --steps 1 --warmup 3 --concurrent 1 --batch 4 --no-cpu-baseline --no-kernel-timing"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_b4.csv $CMD > gpurun_out/ncu.log 2>&1
echo "ncu b4 exit $?"
python tools/launch_summary.py gpurun_out/launches_b4.csv > gpurun_out/launches_b4.summary.txt
head -16 gpurun_out/launches_b4.summary.txt; grep -A24 "by grid" gpurun_out/launches_b4.summary.txt
