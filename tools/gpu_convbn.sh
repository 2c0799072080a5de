#!/bin/bash
# conv tile-width heuristic: probe (auto) + model tests + bench
timeout 300 python tools/conv_probe.py 2>&1 | grep -E "time |PARITY" | cut -c1-100
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_e2vid.py -q -m gpu -x --timeout 300 -p no:cacheprovider 2>&1 | tail -3
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_nocpu.json 2> gpurun_out/bench.err; echo "bench exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_nocpu.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d.get('single_sequence',{}).get('value'), d.get('frame_checksum'))"
