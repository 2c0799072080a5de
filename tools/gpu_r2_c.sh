#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/mlp_probe.py > gpurun_out/mlp_probe.log 2>&1; echo "probe exit $?"; cat gpurun_out/mlp_probe.log
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 300 -p no:cacheprovider -k "mlp" > gpurun_out/pytest_sel.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_sel.log
