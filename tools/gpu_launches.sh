#!/bin/bash
# ncu launch list (per-launch durations) of one short step at batch 4 and batch 1
mkdir -p gpurun_out
CMD="python bench.py --windows 6 --steps 1 --warmup 3 --concurrent 1 --batch 4 --no-cpu-baseline --no-kernel-timing"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_b4.csv $CMD > gpurun_out/ncu.log 2>&1
echo "ncu b4 exit $?"
CMD1="python bench.py --windows 6 --steps 1 --warmup 3 --concurrent 1 --batch 1 --no-cpu-baseline --no-kernel-timing"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_b1.csv $CMD1 > gpurun_out/ncu1.log 2>&1
echo "ncu b1 exit $?"
python tools/launch_summary.py gpurun_out/launches_b4.csv > gpurun_out/launches_b4.summary.txt
python tools/launch_summary.py gpurun_out/launches_b1.csv > gpurun_out/launches_b1.summary.txt
head -24 gpurun_out/launches_b4.summary.txt
