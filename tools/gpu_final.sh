#!/bin/bash
# Round-end style run: full GPU test suite, smoke, both bench arms, and the profile evidence.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm exit $?"
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -2 gpurun_out/bench.err
timeout 900 python bench.py --concurrent 1 --batch 1 --no-cpu-baseline > gpurun_out/bench_single.json 2> gpurun_out/bench_single.err; echo "bench single exit $?"
CMD="python bench.py --windows 6 --steps 1 --warmup 3 --concurrent 1 --batch 4 --no-cpu-baseline --no-kernel-timing"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1
echo "ncu launch list exit $?"
CMD1="python bench.py --windows 4 --steps 1 --warmup 3 --concurrent 1 --batch 1 --no-cpu-baseline --no-kernel-timing"
timeout 600 $CMD1 > gpurun_out/plain1.log 2>&1 && \
timeout 600 ncu --set full --clock-control none -k regex:gemm_tc_kernel -s 3 -c 2 -o gpurun_out/prof_lstm_l1 $CMD1 > gpurun_out/ncu1.log 2>&1; echo "ncu full exit $?"
ncu -i gpurun_out/prof_lstm_l1.ncu-rep --page raw --csv > gpurun_out/prof_lstm_l1.raw.csv 2>/dev/null
ls -la gpurun_out | head -30
