#!/bin/bash
# Targeted `ncu --set full` captures of the implicit-GEMM kernel (small reports; raw CSV extracted on the box).
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
CMD="python bench.py --windows 4 --steps 1 --warmup 3 --concurrent 1 --no-cpu-baseline --no-kernel-timing"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 || exit 1
# gemm launch order for T=4: 0 head, 1-2 enc0 convs, 3-10 LSTM L1 steps, 11.. attention GEMMs, 103.. LSTM L3 steps
timeout 600 ncu --set full --clock-control none -k regex:gemm_tc_kernel -s 3 -c 2 -o gpurun_out/prof_lstm_l1 $CMD > gpurun_out/ncu1.log 2>&1; echo "ncu1 exit $?"
timeout 600 ncu --set full --clock-control none -k regex:gemm_tc_kernel -s 105 -c 2 -o gpurun_out/prof_lstm_l3 $CMD > gpurun_out/ncu2.log 2>&1; echo "ncu2 exit $?"
timeout 600 ncu --set full --clock-control none -k regex:gemm_tc_kernel -s 0 -c 1 -o gpurun_out/prof_head $CMD > gpurun_out/ncu3.log 2>&1; echo "ncu3 exit $?"
for f in prof_lstm_l1 prof_lstm_l3 prof_head; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null
done
ls -la gpurun_out/
