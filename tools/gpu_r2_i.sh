#!/bin/bash
# single-sequence launch list (B = 1): where does a strict batch-1 sequence spend its time?
mkdir -p gpurun_out
CMD="python bench.py --windows 12 --steps 1 --warmup 3 --concurrent 1 --batch 1 --no-cpu-baseline --no-kernel-timing --no-single"
timeout 600 $CMD > gpurun_out/plain_b1.log 2>&1; echo "plain exit $?"; cut -c1-300 gpurun_out/plain_b1.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_b1.csv $CMD > gpurun_out/ncu_b1.log 2>&1
echo "ncu launch list exit $?"
python tools/launch_summary.py gpurun_out/launches_b1.csv > gpurun_out/launches_b1.summary.txt 2>&1; head -30 gpurun_out/launches_b1.summary.txt
