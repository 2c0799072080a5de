#!/bin/bash
# PDL on the attention chain: full GPU tests, then bench with and without the attribute
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_nocpu.json 2> gpurun_out/bench.err; echo "bench exit $?"
BDE2VID_PDL=0 timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_nopdl.json 2> gpurun_out/bench_nopdl.err; echo "bench (no PDL) exit $?"
python -c "
import json
for n in ('bench_nocpu','bench_nopdl'):
    d=json.loads(open('gpurun_out/%s.json'%n).read().strip().splitlines()[-1])
    print(n, d['value'], d['e2e']['value'], d.get('single_sequence',{}).get('value'), d.get('parity_max_abs'), d.get('frame_checksum'))"
