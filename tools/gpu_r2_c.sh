#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/attn_phase_probe.py > gpurun_out/attn_phase.log 2>&1; echo "probe exit $?"; tail -4 gpurun_out/attn_phase.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -q -m gpu --timeout 600 -p no:cacheprovider -k "attention or golden or fused or bench_path" > gpurun_out/pytest_sel.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_sel.log | cut -c1-300
