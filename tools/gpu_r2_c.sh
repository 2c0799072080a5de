#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/attn_tc256_probe.py > gpurun_out/attn_tc256_probe.log 2>&1; echo "probe exit $?"; tail -30 gpurun_out/attn_tc256_probe.log | cut -c1-260
