#!/bin/bash
# source-level ncu capture of the fused MLP kernels (stall sampling per SASS instruction)
mkdir -p gpurun_out
CMD1="python tools/mlp_probe.py"
timeout 300 $CMD1 > gpurun_out/mlp_probe.log 2>&1; echo "probe exit $?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:mlp_fused256_kernel -s 4 -c 1 -o gpurun_out/prof_mlp256_src $CMD1 > gpurun_out/ncu_k1.log 2>&1; echo "ncu mlp256 exit $?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:mlp_fused_kernel -s 4 -c 1 -o gpurun_out/prof_mlp64_src $CMD1 > gpurun_out/ncu_k2.log 2>&1; echo "ncu mlp64 exit $?"
for f in prof_mlp256_src prof_mlp64_src; do ncu -i gpurun_out/$f.ncu-rep --page source --csv > gpurun_out/$f.src.csv 2>/dev/null; ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null; done
rm -f gpurun_out/prof_mlp256_src.ncu-rep gpurun_out/prof_mlp64_src.ncu-rep
ls -la gpurun_out/*src*
