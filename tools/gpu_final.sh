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
CMD1="python bench.py --windows 4 --steps 1 --warmup 3 --concurrent 1 --batch 4 --no-cpu-baseline --no-kernel-timing"
timeout 600 $CMD1 > gpurun_out/plain1.log 2>&1
# ConvLSTM gate conv (L1, then L2, L3 of the same step), level-1 / level-3 fused attention, fused MLPs, voxeliser
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_tma_kernel -s 12 -c 3 -o gpurun_out/prof_conv_lstm $CMD1 > gpurun_out/ncu1.log 2>&1; echo "ncu conv exit $?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn_fused_kernel -s 20 -c 1 -o gpurun_out/prof_attn64 $CMD1 > gpurun_out/ncu2.log 2>&1; echo "ncu attn64 exit $?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn_win256_kernel -s 30 -c 1 -o gpurun_out/prof_attn_win256 $CMD1 > gpurun_out/ncu3.log 2>&1; echo "ncu win256 exit $?"
timeout 600 ncu --set full --clock-control none -k regex:mlp_fused -s 60 -c 2 -o gpurun_out/prof_mlp $CMD1 > gpurun_out/ncu4.log 2>&1; echo "ncu mlp exit $?"
timeout 600 ncu --set full --clock-control none -k regex:voxel_atomic_kernel -s 4 -c 1 -o gpurun_out/prof_voxel $CMD1 > gpurun_out/ncu5.log 2>&1; echo "ncu voxel exit $?"
for f in prof_conv_lstm prof_attn64 prof_attn_win256 prof_mlp prof_voxel; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null; done
ls -la gpurun_out | head -40
