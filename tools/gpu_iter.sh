#!/bin/bash
# Quick iteration run: GPU tests + single-sequence and default bench lines (no CPU baseline).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --concurrent 1 --batch 1 --no-cpu-baseline > gpurun_out/bench_single.json 2> gpurun_out/bench_single.err; echo "bench single exit $?"; tail -3 gpurun_out/bench_single.err
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
for f in ("bench_single", "bench"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value %.1f e2e %.1f ms/step %.1f roof %.3f gemm_ms %.1f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms_per_step"]))
    except Exception as e:
        print(f, "ERR", e)
PY
