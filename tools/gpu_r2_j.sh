#!/bin/bash
# iteration session: MLP kernels (tanh GELU, hoisted LN loads, L2 exchange) -> tests, probe, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_e2vid.py -q -m gpu -x --timeout 300 -p no:cacheprovider > gpurun_out/pytest_sel.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_sel.log
timeout 300 python tools/mlp_probe.py > gpurun_out/mlp_probe.log 2>&1; echo "mlp probe exit $?"; cat gpurun_out/mlp_probe.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_nocpu.json 2> gpurun_out/bench.err; echo "bench exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_nocpu.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d.get('single_sequence',{}).get('value'), d.get('parity_max_abs'))"
