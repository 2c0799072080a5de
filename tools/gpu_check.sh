#!/bin/bash
# First-contact GPU run: each stage under its own timeout so a hung kernel cannot eat the box.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?"; tail -n 25 gpurun_out/$name.log; }
run voxel python -m pytest tests/test_gpu_voxel.py -q -x --timeout 120 -p no:cacheprovider
run simt python -m pytest tests/test_gpu_gemm.py -q --timeout 120 -p no:cacheprovider -k "simt"
run ops python -m pytest tests/test_gpu_ops.py -q --timeout 120 -p no:cacheprovider
run model_fp32 python -m pytest tests/test_gpu_model.py -q --timeout 300 -p no:cacheprovider -k "fp32 or fallback" -s
run tc python -m pytest tests/test_gpu_gemm.py -q --timeout 60 -p no:cacheprovider -k "tcgen05" -s
run model_bf16 python -m pytest tests/test_gpu_model.py -q --timeout 300 -p no:cacheprovider -k "bf16" -s
