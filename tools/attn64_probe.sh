#!/bin/bash
# Builds ablation variants of the level-1 attention kernel (-DBDE_ATTN_PROBE=n, see attn_fused.cu) next to the product library.
# Usage (build container): tools/attn64_probe.sh build ; (GPU box): tools/attn64_probe.sh run
D=bde2vid_b200
if [ "$1" = "build" ]; then
  python -c "from bde2vid_b200 import build; build.build()"
  for n in ${PROBES:-1 2 3 4}; do
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr ${EXTRA} \
      -DBDE_ATTN_PROBE=$n -c $D/csrc/attn_fused.cu -o /tmp/attn_fused_probe$n.o &
  done
  wait
  for n in ${PROBES:-1 2 3 4}; do
    objs=$(ls $D/build/*.o | grep -v attn_fused.o)
    nvcc -shared -o $D/libbde2vid_sm100_probe$n.so $objs /tmp/attn_fused_probe$n.o -lcudart
  done
  ls -la $D/*.so
else
  python tools/attn64_probe.py
  for n in ${PROBES:-1 2 3 4}; do BDE2VID_LIB=$PWD/$D/libbde2vid_sm100_probe$n.so python tools/attn64_probe.py; done
fi
