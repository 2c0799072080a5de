#!/bin/bash
# One process per configuration (a wrong descriptor / barrier protocol must not take the other configurations down).
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; env "$@" timeout 120 python tools/conv_probe.py > gpurun_out/probe_$name.log 2>&1; echo "exit $?"; grep -E "PARITY|FAIL|time|Error|error" gpurun_out/probe_$name.log | tail -25; }
run single BDE2VID_CONV_DUAL=0 PROBE_DBG=1
run dual BDE2VID_CONV_DUAL=1 PROBE_DBG=1
