#!/bin/bash
# tests (selected) + bench + ncu launch list of a 4-sequence T=6 step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -q -m gpu --timeout 600 -p no:cacheprovider -k "attention or golden or fused or bench_path or mlp" > gpurun_out/pytest_sel.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_sel.log | cut -c1-300
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_nocpu.json 2> gpurun_out/bench.err; echo "bench exit $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_nocpu.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['single_sequence'], d['roofline']['frac'], d['roofline_voxeliser']['frac'])
PY
CMD="python bench.py --windows 6 --steps 1 --warmup 3 --concurrent 1 --batch 4 --no-cpu-baseline --no-kernel-timing --no-single"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_b4.csv $CMD > gpurun_out/ncu.log 2>&1
echo "ncu launch list exit $?"
python tools/launch_summary.py gpurun_out/launches_b4.csv > gpurun_out/launches_b4.summary.txt 2>&1; head -30 gpurun_out/launches_b4.summary.txt
