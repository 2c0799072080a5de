#!/bin/bash
# measurement session: voxel probe, the other BASELINE configs, ncu --set full captures of the changed kernels
mkdir -p gpurun_out
timeout 600 python tools/voxel_probe.py > gpurun_out/voxel_probe.log 2>&1; echo "voxel probe exit $?"; cat gpurun_out/voxel_probe.log | cut -c1-200
timeout 900 python bench.py --config e2vid16 --steps 3 > gpurun_out/bench_e2vid16.json 2> gpurun_out/bench_e2vid16.err; echo "e2vid16 exit $?"; cut -c1-400 gpurun_out/bench_e2vid16.json; tail -2 gpurun_out/bench_e2vid16.err
timeout 900 python bench.py --config gen4 --steps 3 > gpurun_out/bench_gen4.json 2> gpurun_out/bench_gen4.err; echo "gen4 exit $?"; cut -c1-400 gpurun_out/bench_gen4.json; tail -2 gpurun_out/bench_gen4.err
timeout 900 python bench.py --config shard64 --steps 2 --warmup 1 > gpurun_out/bench_shard64.json 2> gpurun_out/bench_shard64.err; echo "shard64 exit $?"; cut -c1-400 gpurun_out/bench_shard64.json; tail -2 gpurun_out/bench_shard64.err
timeout 600 ncu --set full --clock-control none -k regex:voxel_atomic_kernel -s 3 -c 3 -o gpurun_out/prof_voxel_gen4 python tools/voxel_probe.py --ncu-shape > gpurun_out/ncu_voxel.log 2>&1; echo "ncu voxel exit $?"
ncu -i gpurun_out/prof_voxel_gen4.ncu-rep --page raw --csv > gpurun_out/prof_voxel_gen4.raw.csv 2>/dev/null
