#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_gemm.py -q --timeout 60 -p no:cacheprovider -x > gpurun_out/gemm.log 2>&1; echo "gemm exit $?"; tail -4 gpurun_out/gemm.log
timeout 300 python -m pytest tests/test_gpu_model.py tests/test_gpu_ops.py -q --timeout 200 -p no:cacheprovider -k "not fp32 or ops" -s > gpurun_out/model.log 2>&1; echo "model exit $?"; grep -E "max-abs|passed|failed|batch2|fused" gpurun_out/model.log | tail -14
for cfg in "$@"; do
  c=${cfg%x*}; b=${cfg#*x}
  timeout 600 python bench.py --steps 3 --warmup 3 --concurrent $c --batch $b --no-cpu-baseline --no-kernel-timing > gpurun_out/bench_${c}x${b}.json 2> gpurun_out/bench_${c}x${b}.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${c}x${b}.json'))
    print('streams=$c batch=$b', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],1), d['clocks'])
except Exception as e: print('ERR $c $b', e)
PY
  tail -2 gpurun_out/bench_${c}x${b}.err
done
