#!/bin/bash
mkdir -p gpurun_out
echo "=== probe"; PROBE_DBG=1 timeout 120 python tools/conv_probe.py > gpurun_out/probe_elect.log 2>&1; echo "exit $?"; grep -E "PARITY|FAIL|time|Error|error" gpurun_out/probe_elect.log | tail -16
bash tools/gpu_iter.sh
