#!/bin/bash
# Per-source-line ncu views (tools/ncu_source_lines.py) of the hot kernels inside a short bench run; CSVs land in gpurun_out/src/
O=gpurun_out/src
mkdir -p $O
CMD1="python bench.py --windows 4 --steps 1 --warmup 3 --concurrent 1 --batch 8 --no-cpu-baseline --no-kernel-timing --no-single"
for spec in "attn_win256_tc_kernel:30:tc256" "mlp_fused256_kernel:30:mlp256" "mlp_fused_kernel:30:mlp64" "conv_tma_kernel:14:conv"; do
  IFS=: read k skip name <<< "$spec"
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$k -s $skip -c 1 -o /tmp/p_$name $CMD1 > $O/ncu_$name.log 2>&1
  ncu -i /tmp/p_$name.ncu-rep --page source --csv --print-source cuda,sass > $O/${name}_src.csv 2>/dev/null
  echo "$name: $(wc -c < $O/${name}_src.csv) bytes"
done
