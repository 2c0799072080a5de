#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_ops.py -q --timeout 60 -p no:cacheprovider -s > gpurun_out/ops.log 2>&1; echo "gemm+ops exit $?"; grep -E "passed|failed|Error|error|chunk-major" gpurun_out/ops.log | tail -12
timeout 300 python -m pytest tests/test_gpu_model.py -q --timeout 200 -p no:cacheprovider -k "bf16" -s > gpurun_out/model.log 2>&1; echo "model exit $?"; grep -E "max-abs|passed|failed|batch2|fused" gpurun_out/model.log | tail -14
for c in $@; do
timeout 600 python bench.py --steps 3 --warmup 3 --concurrent $c --no-cpu-baseline > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err; echo "bench c=$c exit $?"; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_c$c.json'))
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline'] and {k:d['roofline'][k] for k in ('achieved','frac','launches','avg_launch_us','kernel_ms_per_step')}, d['roofline_voxeliser'] and d['roofline_voxeliser']['achieved'], d['clocks'])
except Exception as e: print('ERR', e)
PY
tail -3 gpurun_out/bench_c$c.err
done
CMD="python bench.py --windows 6 --steps 1 --warmup 3 --concurrent 1 --no-cpu-baseline --no-kernel-timing"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 6000 --csv --log-file gpurun_out/launches_warm.csv $CMD > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"
