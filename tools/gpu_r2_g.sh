#!/bin/bash
# evidence session: ncu --set full captures of the round-2 kernels + voxeliser DRAM traffic without cache flushes
mkdir -p gpurun_out
CMD1="python bench.py --windows 4 --steps 1 --warmup 3 --concurrent 1 --batch 4 --no-cpu-baseline --no-kernel-timing --no-single"
timeout 600 $CMD1 > gpurun_out/plain1.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn_win256_tc_kernel -s 30 -c 1 -o gpurun_out/prof_attn_tc256 $CMD1 > gpurun_out/ncu1.log 2>&1; echo "ncu tc256 exit $?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn_fused_kernel -s 20 -c 1 -o gpurun_out/prof_attn64 $CMD1 > gpurun_out/ncu2.log 2>&1; echo "ncu attn64 exit $?"
timeout 600 ncu --set full --clock-control none -k regex:mlp_fused256_kernel -s 30 -c 1 -o gpurun_out/prof_mlp256 $CMD1 > gpurun_out/ncu3.log 2>&1; echo "ncu mlp256 exit $?"
timeout 600 ncu --set full --clock-control none -k regex:conv_tma_kernel -s 12 -c 2 -o gpurun_out/prof_conv_lstm $CMD1 > gpurun_out/ncu4.log 2>&1; echo "ncu conv exit $?"
for f in prof_attn_tc256 prof_attn64 prof_mlp256 prof_conv_lstm; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null; done
# voxeliser: every launch of one call (memset nodes + reduction kernels), caches NOT flushed between launches
cat > /tmp/vox_one.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from bde2vid_b200 import ops, synth
for (H, W, N, T, Hp, Wp, pt, pl) in ((260, 346, 31500, 100, 264, 352, 2, 3), (720, 1280, 333333, 64, 720, 1280, 0, 0)):
    ev = synth.gen_events(0, T, H, W, N)
    f32 = [torch.from_numpy(a).cuda() for a in synth.to_loader_format_seq(ev)]
    out = torch.empty(T, 5, Hp, Wp, device='cuda')
    for _ in range(2):
        ops.voxelize_seq(*f32, 5, H, W, pt, pl, Hp, Wp, out=out, min_events=3)
    torch.cuda.synchronize()
PY
timeout 600 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/voxel_traffic.csv python /tmp/vox_one.py > gpurun_out/ncu5.log 2>&1; echo "ncu voxel traffic exit $?"
ls -la gpurun_out/*.ncu-rep | head
