#!/bin/bash
# round 2, session B: full GPU test suite (no -x) + the other BASELINE configs, kept short
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -40 gpurun_out/pytest_gpu.log | cut -c1-300
