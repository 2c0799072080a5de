#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
CMD="python bench.py --windows 4 --steps 1 --warmup 3 --concurrent 1 --no-cpu-baseline --no-kernel-timing"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 || exit 1
# gemm launch order T=4: 0 head, 1-2 enc0, 3-10 LSTM L1, 11.. attention L1: per block qproj(190,1) kv(568,1) proj(190,1) fc1(182,2) fc2(182,1)
timeout 600 ncu --set full --import-source on --clock-control none --cache-control none -k regex:gemm_tc_kernel -s 11 -c 5 -o gpurun_out/prof_attn_l1 $CMD > gpurun_out/ncu1.log 2>&1; echo "ncu1 exit $?"
# L3 attention gemms start after: 11 + 80 (L1 attn) + 2 + 8 + 2 + 8 = 111
timeout 600 ncu --set full --import-source on --clock-control none --cache-control none -k regex:gemm_tc_kernel -s 111 -c 5 -o gpurun_out/prof_attn_l3 $CMD > gpurun_out/ncu2.log 2>&1; echo "ncu2 exit $?"
for f in prof_attn_l1 prof_attn_l3; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null
  ncu -i gpurun_out/$f.ncu-rep --page source --csv > gpurun_out/$f.source.csv 2>/dev/null
done
ls -la gpurun_out/ | head -20
